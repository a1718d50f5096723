"""CPU oracle for the ukbb_cardiac FCN deploy hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``ukbb_cardiac_b200/`` or
``common/`` may import this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs do, and there only as the checker or as the timed CPU arm.

PARITY UNPINNED for the network part: the reference ships no tests, golden
vectors or fixtures and its arithmetic lives in
TensorFlow 1.x (version unpinned upstream, not installable here), so the
network oracle is a restatement of the published TF op semantics anchored on
the reference's call sites (``common/network.py:19-25,117-230``,
``common/train_network.py:142-199``, ``common/deploy_network.py:43-225``).
The preprocessing part IS pinned: ``tests/golden/rescale_*.npz`` were produced
by the reference's own ``rescale_intensity`` (``common/image_utils.py:70-77``)
imported in the build container (see ``tests/golden/make_golden.py``).
"""

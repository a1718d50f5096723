"""Host-side engine: the Python mirror of the reference's operator boundary.

The reference drives the network through
``sess.run(['prob:0', 'pred:0'], feed_dict={'image:0': image, 'training:0': False})``
(``common/deploy_network.py:110-111, 195-196``) after restoring the graph with
``tf.train.import_meta_graph`` / ``saver.restore`` (``:48-49``).  ``FCNEngine`` keeps
those names and argument meanings (``Session``-style ``run``), and adds the
whole-volume calls the B200 design is built around (one call per subject
instead of one per frame).  PyTorch is used only to own device / pinned
memory and streams; every kernel lives in libukbb_fcn.so.
"""
from __future__ import annotations

import ctypes as C
import math
import threading
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, tf_bundle
from . import weights as W

MODES = {"fp32": _lib.MODE_FP32, "bf16": _lib.MODE_BF16, "fp16": _lib.MODE_FP16,
         "bf16x3": _lib.MODE_BF16X3, "fp16x3": _lib.MODE_FP16X3, "fp16x2": _lib.MODE_FP16X2}
# The default is the split-operand FP16 tensor-core mode with FP8 correction products ("x2", include/ukbb_fcn.h): the fastest mode that
# meets the parity tolerance (>= 99.9 % label agreement, Dice >= 0.999 against the float32 reference on random-init weights; DESIGN.md
# section 5).  "fp16x3" (three FP16 products per K step) is ~9 % slower and ~4x tighter on the logits.
# Plain "bf16" / "fp16" are faster but do NOT meet it; "fp32" is the CUDA-core exactness mode.
DEFAULT_MODE = "fp16x2"


def pad16(x: int) -> Tuple[int, int]:
    """deploy_network.py:97-98 -> (X2, x_pre)."""
    x2 = int(math.ceil(x / 16.0)) * 16
    return x2, int((x2 - x) / 2)


def _fptr(a: Optional[np.ndarray]):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(C.c_float))


class FCNEngine:
    """build_FCN inference engine bound to one CUDA device."""

    def __init__(self, tensors: Dict[str, np.ndarray], device: int = 0, mode: str = DEFAULT_MODE):
        if mode not in MODES:
            raise ValueError("mode must be one of %s" % sorted(MODES))
        if not torch.cuda.is_available():
            raise RuntimeError("FCNEngine needs a CUDA device: libukbb_fcn has no CPU fallback")
        self.lib = _lib.load()
        self.n_class = W.validate(tensors)
        self.mode = mode
        self.device = torch.device("cuda", device)
        tab = W.layer_table(self.n_class)
        keep = []          # keep host arrays alive during the create call
        convs = (_lib.ConvWeights * W.N_CONV)()
        for i, sp in enumerate(tab):
            k = np.ascontiguousarray(tensors[W.conv_name(i) + "/kernel"], dtype=np.float32)
            keep.append(k)
            cw = convs[i]
            cw.kernel = _fptr(k)
            cw.ksize, cw.cin, cw.cout, cw.stride = sp.ksize, sp.cin, sp.cout, sp.stride
            if i < W.N_BN:
                arrs = [np.ascontiguousarray(tensors[W.bn_name(i) + "/" + v], dtype=np.float32)
                        for v in ("gamma", "beta", "moving_mean", "moving_variance")]
                keep.extend(arrs)
                cw.gamma, cw.beta, cw.moving_mean, cw.moving_variance = (_fptr(a) for a in arrs)
            else:
                b = np.ascontiguousarray(tensors[W.conv_name(i) + "/bias"], dtype=np.float32)
                keep.append(b)
                cw.bias = _fptr(b)
        fw = _lib.FcnWeights(W.N_CONV, convs, W.BN_EPS)
        handle = C.c_void_p()
        _lib.check(self.lib.ukbb_fcn_create(C.byref(fw), self.n_class, device, MODES[mode], C.byref(handle)))
        self._h = handle
        self._slot = 0
        self._pin_pool, self._pin_live, self._pin_lock = [], {}, threading.Lock()
        self._join_stream = None

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_checkpoint(cls, model_path: str, device: int = 0, mode: str = DEFAULT_MODE) -> "FCNEngine":
        """``saver.restore(sess, model_path)``: reads ``model_path.index`` / ``.data-*``."""
        _lib.load()
        tensors = tf_bundle.read_bundle(model_path)
        return cls(tensors, device=device, mode=mode)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.ukbb_fcn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ device-level calls
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def forward(self, image: torch.Tensor, x_pre: int = 0, y_pre: int = 0, x: Optional[int] = None,
                y: Optional[int] = None, want_logits: bool = False, want_prob: bool = False):
        """image: cuda float32 [N, Y2, X2] (rescaled, padded).  Returns (labels uint8 [N, Y, X],
        logits or None, prob or None) with logits/prob as [N, Y2, X2, C]."""
        assert image.is_cuda and image.dtype == torch.float32 and image.dim() == 3 and image.is_contiguous()
        n, y2, x2 = image.shape
        x = x2 - x_pre if x is None else x
        y = y2 - y_pre if y is None else y
        labels = torch.empty((n, y, x), dtype=torch.uint8, device=self.device)
        logits = torch.empty((n, y2, x2, self.n_class), dtype=torch.float32, device=self.device) if want_logits else None
        prob = torch.empty((n, y2, x2, self.n_class), dtype=torch.float32, device=self.device) if want_prob else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ukbb_fcn_forward(
                self._h, image.data_ptr(), n, x2, y2, x_pre, y_pre, x, y, labels.data_ptr(),
                logits.data_ptr() if want_logits else None, prob.data_ptr() if want_prob else None, self._stream()))
        return labels, logits, prob

    def class_counts(self, n: int) -> torch.Tensor:
        out = torch.empty((n, self.n_class), dtype=torch.int64, device=self.device)
        _lib.check(self.lib.ukbb_fcn_class_counts(self._h, out.data_ptr(), n, self._stream()))
        return out

    def preprocess(self, vol: torch.Tensor, n_slices: int, x: int, y: int, q: Sequence[float] = (1.0, 99.0),
                   clip_in_place: bool = False, out: Optional[torch.Tensor] = None, vlvh: Optional[torch.Tensor] = None):
        """vol: cuda float32 with n_slices*y*x voxels in NIfTI order.  Returns (padded [N, Y2, X2]
        float32, vl_vh cuda float64[2], (x_pre, y_pre)).  `out` / `vlvh` may be passed in to reuse buffers (e.g. when the
        call runs on a side stream so that it overlaps the forward of the previous subject)."""
        assert vol.is_cuda and vol.dtype == torch.float32 and vol.is_contiguous() and vol.numel() == n_slices * x * y
        x2, x_pre = pad16(x)
        y2, y_pre = pad16(y)
        if out is None:
            out = torch.empty((n_slices, y2, x2), dtype=torch.float32, device=self.device)
        if vlvh is None:
            vlvh = torch.empty(2, dtype=torch.float64, device=self.device)
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == n_slices * y2 * x2
        assert vlvh.is_cuda and vlvh.dtype == torch.float64 and vlvh.numel() >= 2
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ukbb_fcn_preprocess(
                self._h, vol.data_ptr(), n_slices, x, y, float(q[0]), float(q[1]), x2, y2, x_pre, y_pre,
                out.data_ptr(), vlvh.data_ptr(), int(clip_in_place), self._stream()))
        return out, vlvh, (x_pre, y_pre)

    def rescale(self, vol: torch.Tensor, n_slices: int, x: int, y: int, vl: float, vh: float,
                clip_in_place: bool = False, out: Optional[torch.Tensor] = None):
        """Rescale + pad a block of slices with thresholds (vl, vh) that were taken over the whole sequence elsewhere
        (`split_blocks` / `SplitEngine`).  vol: cuda float32, n_slices*y*x voxels.  Returns (padded [N, Y2, X2], (x_pre, y_pre))."""
        assert vol.is_cuda and vol.dtype == torch.float32 and vol.is_contiguous() and vol.numel() == n_slices * x * y
        x2, x_pre = pad16(x)
        y2, y_pre = pad16(y)
        if out is None:
            out = torch.empty((n_slices, y2, x2), dtype=torch.float32, device=self.device)
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == n_slices * y2 * x2
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ukbb_fcn_rescale(self._h, vol.data_ptr(), n_slices, x, y, float(vl), float(vh), x2, y2, x_pre, y_pre,
                                                 out.data_ptr(), int(clip_in_place), self._stream()))
        return out, (x_pre, y_pre)

    # ------------------------------------------------------------------ host-level calls
    def segment_host_async(self, vol: torch.Tensor, shape: Tuple[int, int, int, int], labels: torch.Tensor,
                           vl_vh: Optional[torch.Tensor] = None, counts: Optional[torch.Tensor] = None,
                           q: Sequence[float] = (1.0, 99.0)) -> None:
        """Whole-subject call on (pinned) HOST tensors: vol float32 with X*Y*Z*T voxels in NIfTI
        order, labels uint8 of the same size.  Asynchronous; alternate staging slots are used so
        consecutive subjects overlap H2D / compute / D2H.  Call ``sync()`` before reading."""
        x, y, z, t = shape
        assert not vol.is_cuda and vol.dtype == torch.float32 and vol.is_contiguous() and vol.numel() == x * y * z * t
        assert not labels.is_cuda and labels.dtype == torch.uint8 and labels.is_contiguous() and labels.numel() == vol.numel()
        if vl_vh is not None:
            assert vl_vh.dtype == torch.float64 and vl_vh.numel() >= 2 and not vl_vh.is_cuda
        if counts is not None:
            assert counts.dtype == torch.int64 and counts.numel() >= z * t * self.n_class and not counts.is_cuda
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ukbb_fcn_segment_host(
                self._h, vol.data_ptr(), x, y, z, t, float(q[0]), float(q[1]), labels.data_ptr(),
                vl_vh.data_ptr() if vl_vh is not None else None,
                counts.data_ptr() if counts is not None else None, self._slot, self._stream()))
        self._slot ^= 1

    def join(self) -> None:
        _lib.check(self.lib.ukbb_fcn_join(self._h, self._stream()))

    def sync(self) -> None:
        _lib.check(self.lib.ukbb_fcn_sync(self._h))

    # ---- pipelined whole-volume calls (deploy.py): decode of subject i + 1, device work of subject i and encode of subject i - 1 overlap
    def host_buffer(self, nbytes: int) -> np.ndarray:
        """A uint8 numpy view of pinned host memory from a small reuse pool (cudaHostAlloc is slow: ~10 ms per 80 MB).
        Thread-safe: deploy.py's reader threads decode into these buffers."""
        with self._pin_lock:
            best = -1
            for i, t in enumerate(self._pin_pool):
                if t.numel() >= nbytes and (best < 0 or t.numel() < self._pin_pool[best].numel()):
                    best = i
            buf = self._pin_pool.pop(best) if best >= 0 else None
        if buf is None:
            buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=True)
        arr = buf.numpy()
        with self._pin_lock:
            self._pin_live[arr.ctypes.data] = buf
        return arr

    def release_host_buffer(self, arr: np.ndarray) -> None:
        """Give a `host_buffer` array (the array itself, not a view of it) back to the pool."""
        with self._pin_lock:
            t = self._pin_live.pop(arr.ctypes.data, None)
            if t is not None:
                self._pin_pool.append(t)

    def submit_volume(self, image: np.ndarray, q: Sequence[float] = (1.0, 99.0)):
        """Asynchronous `segment_volume`: image (X, Y[, Z[, T]]) float32; when it is a Fortran-ordered view of a `host_buffer` it is
        uploaded from where it lies.  Returns a ticket for `collect_volume`."""
        shp = tuple(image.shape)
        x, y = shp[0], shp[1]
        z = shp[2] if image.ndim >= 3 else 1
        t = shp[3] if image.ndim == 4 else 1
        if image.ndim not in (2, 3, 4):
            raise ValueError("expected a 2-D, 3-D or 4-D image, got shape %s" % (shp,))
        n = x * y * z * t
        own = None
        if image.dtype == np.float32 and image.flags.f_contiguous:
            flat = image.reshape(-1, order="F")                       # a view: NIfTI memory order = Fortran order
        else:
            own = self.host_buffer(n * 4)
            flat = own[:n * 4].view(np.float32)
            flat[:] = np.asarray(image, dtype=np.float32).reshape(-1, order="F")
        lab_buf = self.host_buffer(n)
        aux = self.host_buffer(16 + z * t * self.n_class * 8)
        vol_t = torch.from_numpy(flat)
        lab_t = torch.from_numpy(lab_buf[:n])
        vlvh_t = torch.from_numpy(aux[:16].view(np.float64))
        counts_t = torch.from_numpy(aux[16:16 + z * t * self.n_class * 8].view(np.int64))
        self.segment_host_async(vol_t, (x, y, z, t), lab_t, vlvh_t, counts_t, q)
        if self._join_stream is None:
            self._join_stream = torch.cuda.Stream(self.device)
        side = self._join_stream
        ev = torch.cuda.Event()
        with torch.cuda.stream(side):                                 # the read-back is awaited on a side stream: the next forward is not held up
            self.join()
            ev.record(side)
        return {"ev": ev, "shape": shp, "n": n, "zt": (z, t), "lab": lab_buf, "aux": aux, "own": own, "keep": (vol_t, lab_t, vlvh_t, counts_t)}

    def collect_volume(self, ticket):
        """Wait for a `submit_volume` ticket: (labels uint8 view of pinned memory, Fortran order, (vl, vh), counts int64 [T, Z, C]).
        The labels stay valid until `release_ticket`."""
        ticket["ev"].synchronize()
        n, (z, t) = ticket["n"], ticket["zt"]
        lab = ticket["lab"][:n].reshape(ticket["shape"], order="F")
        vlvh = ticket["aux"][:16].view(np.float64)
        counts = ticket["aux"][16:16 + z * t * self.n_class * 8].view(np.int64).reshape(t, z, self.n_class).copy()
        return lab, (float(vlvh[0]), float(vlvh[1])), counts

    def release_ticket(self, ticket) -> None:
        for k in ("lab", "aux", "own"):
            if ticket.get(k) is not None:
                self.release_host_buffer(ticket[k])
                ticket[k] = None
        ticket["keep"] = None

    def segment_rescaled(self, image2: np.ndarray):
        """Segment an ALREADY rescaled image (X, Y[, Z[, T]]) -- the array `rescale_intensity` returned, before the padding of
        deploy_network.py:97-100 -- without the device percentile pass.  Used for inputs whose native dtype is not float32, where
        the reference's in-place clipping truncates the thresholds to that dtype (deploy.py).  Returns (labels uint8, same shape,
        Fortran order; counts int64 [T, Z, n_class])."""
        shp = tuple(image2.shape)
        x, y = shp[0], shp[1]
        z = shp[2] if image2.ndim >= 3 else 1
        t = shp[3] if image2.ndim == 4 else 1
        x2, x_pre = pad16(x)
        y2, y_pre = pad16(y)
        vol = np.asarray(image2, dtype=np.float32).reshape((x, y, z * t), order="F")
        padded = np.zeros((z * t, y2, x2), dtype=np.float32)
        padded[:, y_pre:y_pre + y, x_pre:x_pre + x] = np.transpose(vol, (2, 1, 0))
        dev = torch.from_numpy(padded).to(self.device)
        labels, _, _ = self.forward(dev, x_pre, y_pre, x, y)
        counts = self.class_counts(z * t).cpu().numpy().reshape(t, z, self.n_class)
        lab = np.transpose(labels.cpu().numpy(), (2, 1, 0)).reshape(shp, order="F")
        return np.asfortranarray(lab), counts

    def segment_volume(self, image: np.ndarray, q: Sequence[float] = (1.0, 99.0)):
        """image: (X, Y, Z, T), (X, Y, Z) or (X, Y) array as returned by ``nim.get_data()``.
        Returns (labels uint8 array of the same shape, Fortran order, (vl, vh), counts
        int64 [T, Z, n_class])."""
        shp = tuple(image.shape)
        if image.ndim == 2:
            x, y, z, t = shp[0], shp[1], 1, 1
        elif image.ndim == 3:
            x, y, z, t = shp[0], shp[1], shp[2], 1
        elif image.ndim == 4:
            x, y, z, t = shp
        else:
            raise ValueError("expected a 2-D, 3-D or 4-D image, got shape %s" % (shp,))
        n = x * y * z * t
        vol = torch.empty(n, dtype=torch.float32, pin_memory=True)
        # NIfTI memory order = Fortran order of the (X, Y, Z, T) array
        vol.numpy()[:] = np.asarray(image, dtype=np.float32).reshape(-1, order="F")
        labels = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        vlvh = torch.empty(2, dtype=torch.float64, pin_memory=True)
        counts = torch.empty((t, z, self.n_class), dtype=torch.int64, pin_memory=True)
        self.segment_host_async(vol, (x, y, z, t), labels, vlvh, counts, q)
        self.sync()
        lab = labels.numpy().reshape(shp, order="F").copy(order="F")
        return lab, (float(vlvh[0]), float(vlvh[1])), counts.numpy().copy()

    # ------------------------------------------------------------------ reference operator API
    def run(self, fetches, feed_dict):
        """Drop-in for ``sess.run(['prob:0', 'pred:0'], feed_dict={'image:0': image_NXYC,
        'training:0': False})``: returns the fetched arrays in TF's layouts
        (prob float32 [N, X, Y, C], pred int32 [N, X, Y])."""
        single = isinstance(fetches, str)
        names = [fetches] if single else list(fetches)
        for nm in names:
            if nm not in ("prob:0", "pred:0", "logits:0"):
                raise KeyError("unknown fetch %r (the deploy graph exposes prob:0 and pred:0)" % nm)
        if "image:0" not in feed_dict:
            raise KeyError("feed_dict must provide 'image:0'")
        if feed_dict.get("training:0", False):
            raise NotImplementedError("training:0=True (batch statistics) is not part of the deploy path")
        img = np.asarray(feed_dict["image:0"], dtype=np.float32)
        if img.ndim != 4 or img.shape[3] != 1:
            raise ValueError("image:0 must have shape [N, X, Y, 1], got %s" % (img.shape,))
        n, xx, yy, _ = img.shape
        if xx % 16 or yy % 16:
            raise ValueError("image:0 spatial size %dx%d must be a multiple of 16 "
                             "(deploy_network.py:97-100 pads before calling)" % (xx, yy))
        dev = torch.from_numpy(np.ascontiguousarray(img[..., 0])).to(self.device)      # [N, X, Y]
        dev = dev.permute(0, 2, 1).contiguous()                                          # [N, Y, X]
        labels, logits, prob = self.forward(dev, want_logits="logits:0" in names, want_prob="prob:0" in names)
        out = []
        for nm in names:
            if nm == "pred:0":
                out.append(labels.permute(0, 2, 1).to(torch.int32).cpu().numpy())
            elif nm == "prob:0":
                out.append(prob.permute(0, 2, 1, 3).contiguous().cpu().numpy())
            else:
                out.append(logits.permute(0, 2, 1, 3).contiguous().cpu().numpy())
        return out[0] if single else out

    @property
    def split(self) -> bool:
        return self.mode in ("bf16x3", "fp16x3", "fp16x2")

    @property
    def dtype16(self) -> torch.dtype:
        return torch.float16 if self.mode in ("fp16", "fp16x3", "fp16x2") else torch.bfloat16

    def debug_conv(self, layer: int, x: torch.Tensor, level_out: int) -> torch.Tensor:
        """Test hook: one tensor-core conv layer on a cuda 16-bit [N, H, W, Cin] tensor (rows = Y); in the x3 modes
        x and the result are [2, N, H, W, C] = hi plane, lo plane."""
        sp = W.layer_table(self.n_class)[layer]
        dt = self.dtype16
        assert x.is_cuda and x.dtype == dt and x.is_contiguous() and x.shape[-1] == sp.cin
        assert x.dim() == (5 if self.split else 4) and (not self.split or x.shape[0] == 2)
        n, hi, wi, _ = x.shape[-4:]
        ho, wo = -(-hi // sp.stride), -(-wi // sp.stride)
        shape = (n, ho, wo, sp.cout)
        out = torch.empty(((2,) + shape) if self.split else shape, dtype=dt, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ukbb_fcn_debug_conv(self._h, layer, x.data_ptr(), n, hi, wi, level_out,
                                                    out.data_ptr(), self._stream()))
        return out

    def debug_read(self, which: int, level: int, shape: Tuple[int, ...]) -> torch.Tensor:
        """Test hook: an intermediate tensor of the most recent forward as float32 (hi + lo in the x3 modes).
        which: 0 / 1 = encoder ping / pong buffer of `level`, 2 = t_level; shape = (N, H_l, W_l, C) of this call."""
        out = torch.empty(shape, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ukbb_fcn_debug_read(self._h, which, level, out.data_ptr(), out.numel(), self._stream()))
        return out

    def debug_flags(self, flags: int) -> None:
        """Test hook: bit 0 = always take the generic radix select in the percentile rescale (no integer fast path)."""
        _lib.check(self.lib.ukbb_fcn_debug_flags(self._h, int(flags)))

    def kernel_timer(self, enable: bool) -> None:
        """Bracket every launch of the fused head kernel with CUDA events on the launching stream (bench.py roofline)."""
        _lib.check(self.lib.ukbb_fcn_kernel_timer(self._h, 1 if enable else 0))

    def kernel_timer_read(self):
        """(summed kernel time in ms, launches) since the last read; synchronises the device."""
        import ctypes as C
        ms, n = C.c_double(0.0), C.c_longlong(0)
        _lib.check(self.lib.ukbb_fcn_kernel_timer_read(self._h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    @property
    def launch_count(self) -> int:
        return int(self.lib.ukbb_fcn_launch_count(self._h))


def split_blocks(n: int, parts: int):
    """n slices -> `parts` contiguous blocks [(start, stop), ...], sizes differing by at most one (empty blocks when parts > n)."""
    base, extra = divmod(n, parts)
    out, a = [], 0
    for g in range(parts):
        b = a + base + (1 if g < extra else 0)
        out.append((a, b))
        a = b
    return out


class SplitEngine:
    """ONE sequence over several GPUs (single-subject latency with G > 1), one process driving all devices.

    The only coupling between the slices of a sequence is the pair of percentile thresholds of rescale_intensity
    (common/image_utils.py:70-77, taken over the whole 4-D array at deploy_network.py:89).  GPU 0 receives the whole volume, takes
    the thresholds (and keeps the rescaled slices of its own block); the other GPUs receive only their contiguous block of (z, t)
    slices while that runs, rescale it with the 16 bytes (vl, vh) the host hands over, and every GPU runs the network on its block.
    No device collective; labels and class counts are bit-identical to one FCNEngine.segment_volume call."""

    def __init__(self, weights: Dict[str, np.ndarray], devices: Sequence[int], mode: str = DEFAULT_MODE):
        assert len(devices) >= 1
        self.devices = [int(d) for d in devices]
        self.engines = [FCNEngine(weights, device=d, mode=mode) for d in self.devices]
        self.n_class = self.engines[0].n_class

    def close(self) -> None:
        for e in self.engines:
            e.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def segment_volume(self, image: np.ndarray, q: Sequence[float] = (1.0, 99.0)):
        """image: (X, Y, Z, T), (X, Y, Z) or (X, Y).  Returns what FCNEngine.segment_volume returns."""
        shp = tuple(image.shape)
        x, y = shp[0], shp[1]
        z = shp[2] if image.ndim >= 3 else 1
        t = shp[3] if image.ndim == 4 else 1
        n, sl = z * t, x * y
        host = torch.empty(n * sl, dtype=torch.float32, pin_memory=True)
        host.numpy()[:] = np.asarray(image, dtype=np.float32).reshape(-1, order="F")
        blocks = split_blocks(n, len(self.engines))
        e0 = self.engines[0]
        # blocks of the other GPUs first (asynchronous copies), then the whole volume to GPU 0
        dev_blocks = [None] * len(self.engines)
        for g, (a, b) in enumerate(blocks):
            if g > 0 and b > a:
                with torch.cuda.device(self.devices[g]):
                    dev_blocks[g] = host[a * sl:b * sl].to(self.engines[g].device, non_blocking=True)
        with torch.cuda.device(self.devices[0]):
            vol0 = host.to(e0.device, non_blocking=True)
            padded0, vlvh_dev, (x_pre, y_pre) = e0.preprocess(vol0, n, x, y, q)
            vlvh = vlvh_dev.cpu()                                    # synchronises GPU 0: the 16 bytes every other GPU needs
        vl, vh = float(vlvh[0]), float(vlvh[1])
        labels = torch.empty((n, y, x), dtype=torch.uint8, pin_memory=True)
        counts = torch.empty((n, self.n_class), dtype=torch.int64, pin_memory=True)
        for g, (a, b) in enumerate(blocks):
            if b == a:
                continue
            e = self.engines[g]
            with torch.cuda.device(self.devices[g]):
                if g == 0:
                    padded = padded0[a:b]
                else:
                    padded, _ = e.rescale(dev_blocks[g], b - a, x, y, vl, vh)
                lab, _, _ = e.forward(padded, x_pre, y_pre, x, y)
                labels[a:b].copy_(lab, non_blocking=True)
                counts[a:b].copy_(e.class_counts(b - a), non_blocking=True)
        for g, (a, b) in enumerate(blocks):
            if b > a:
                torch.cuda.synchronize(self.devices[g])
        lab = labels.numpy().reshape(-1).reshape(shp, order="F").copy(order="F")      # [n][y][x] is the NIfTI order of the volume
        return lab, (vl, vh), counts.numpy().reshape(t, z, self.n_class).copy()

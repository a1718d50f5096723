#!/usr/bin/env python3
"""Drop-in for the reference's ``common/deploy_network.py`` (same flags, files and prints):

    python3 common/deploy_network.py --seq_name sa --data_dir demo_image --model_path trained_model/FCN_sa

(the invocations of demo_pipeline.py:63-64, 89-96).  The implementation lives in
``ukbb_cardiac_b200/deploy.py``; the arithmetic runs in libukbb_fcn.so on a B200.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from ukbb_cardiac_b200.deploy import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main())

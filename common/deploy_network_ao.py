#!/usr/bin/env python3
"""Drop-in for the reference's ``common/deploy_network_ao.py`` (UNet-LSTM aortic model): same command line, same files, same
output lines; the implementation is ``ukbb_cardiac_b200/deploy_ao.py`` on top of libukbb_fcn.so (sm_100a)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from ukbb_cardiac_b200.deploy_ao import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main())

"""`import gru_vae` shim: put this directory ahead of the reference's src/nets on PYTHONPATH
(egs/one-to-one/path.sh:11) and the reference's train / decode scripts pick up the B200 path."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from cyclevae_vc_b200.gru_vae import *  # noqa: F401,F403,E402
from cyclevae_vc_b200.gru_vae import GRU_RNN, TWFSEloss, initialize, loss_vae, sampling_vae_batch  # noqa: F401,E402

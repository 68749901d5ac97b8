"""
Constants and dtype helpers, mirroring src/vegasflow/configflow.py of the
reference (file:line citations are relative to /root/reference).

fp64 is the only arithmetic the CUDA path implements; VEGASFLOW_FLOAT=32 is
accepted with a warning and ignored.
"""
import logging
import os

import torch

# configflow.py:13-24
BINS_MAX = 50
ALPHA = 1.5
BETA = 0.75  # Vegas+
TECH_CUT = 1e-8
MAX_EVENTS_LIMIT = int(1e6)
MAX_NEVAL_HCUBE = int(1e4)

# configflow.py:27 -- kept for signature compatibility; one process drives one GPU
DEFAULT_ACTIVE_DEVICES = ["GPU"]

# configflow.py:30-61
LOG_DICT = {"0": logging.ERROR, "1": logging.WARNING, "2": logging.INFO, "3": logging.DEBUG}
_log_level_idx = os.environ.get("VEGASFLOW_LOG_LEVEL")
_float_env = os.environ.get("VEGASFLOW_FLOAT", "64")
_int_env = os.environ.get("VEGASFLOW_INT", "32")

_log_level = LOG_DICT.get(_log_level_idx, LOG_DICT["2"])
logger = logging.getLogger(__name__.split(".")[0])
logger.setLevel(_log_level)
if not logger.handlers:
    _console_handler = logging.StreamHandler()
    _console_handler.setLevel(_log_level)
    _console_handler.setFormatter(logging.Formatter("[%(levelname)s] (%(name)s) %(message)s"))
    logger.addHandler(_console_handler)

# configflow.py:69-86
DTYPE = torch.float64
if _float_env != "64":
    logger.warning("VEGASFLOW_FLOAT=%s: the B200 path computes in float64 only", _float_env)
if _int_env == "64":
    DTYPEINT = torch.int64
else:
    DTYPEINT = torch.int32
    if _int_env != "32":
        logger.warning("VEGASFLOW_INT=%s not understood, defaulting to 32 bits", _int_env)

FMAX = torch.finfo(torch.float64).max


def run_eager(flag=True):
    """configflow.py:89-96.  Nothing is traced here; kept as a no-op."""
    return None


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None


def int_me(i):
    """Cast the input to DTYPEINT (configflow.py:102-104)."""
    if isinstance(i, torch.Tensor):
        return i.to(DTYPEINT)
    return torch.as_tensor(i, dtype=DTYPEINT, device=_device())


def float_me(i):
    """Cast the input to DTYPE (configflow.py:107-109)."""
    if isinstance(i, torch.Tensor):
        return i.to(DTYPE)
    return torch.as_tensor(i, dtype=DTYPE, device=_device())

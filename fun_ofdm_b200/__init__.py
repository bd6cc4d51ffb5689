"""fun_ofdm_b200 — B200-native batched 802.11a receive hot path behind fun_ofdm's receiver API.

The product is the C-ABI library ``fun_ofdm_b200/lib/libb200rx.so`` (hand-written sm_100a CUDA,
``include/b200rx.h``) and the C++ host adapter in ``fun_ofdm_b200/host`` that mirrors the
reference's ``fun::block`` / ``receiver_chain`` interface.  This Python package is the thin binding
used by the tests and ``bench.py``: ctypes over the C ABI, with PyTorch supplying device memory,
streams and ``torch.distributed`` only.  There is no CPU implementation of the path here.
"""
from .rx import (  # noqa: F401
    B200RxError,
    Receiver,
    FMT_FC64, FMT_FC32, FMT_SC16, FMT_TAGGED_FC64,
    ST_OK, ST_HDR_PARITY, ST_HDR_RATE, ST_CRC_FAIL, ST_TRUNCATED, ST_TOO_LONG,
    RATE_PARAMS, lib_path, load_library, num_symbols, window_samples,
)

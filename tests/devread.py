"""Reads symbol blocks of a panel's TILED device planes ([block of 32 symbols][bar][32 symbols])
straight from HBM -- for parity checks of panels too large to download whole (50,000 x 5,040 is 42 GB of results)."""
import ctypes as C

import numpy as np

_rt = None


def _cudart():
    global _rt
    if _rt is None:
        import torch  # noqa: F401  (loads the CUDA runtime the engine also uses)
        for name in ("libcudart.so.12", "libcudart.so"):
            try:
                _rt = C.CDLL(name)
                break
            except OSError:
                continue
        if _rt is None:
            import glob
            import os
            cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
            _rt = C.CDLL(cands[0])
        _rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        _rt.cudaMemcpy.restype = C.c_int
    return _rt


def read_block(dev_ptr: int, block: int, bars_padded: int, n_bars: int) -> np.ndarray:
    """-> float64 [32 symbols, n_bars] of symbol block `block` of the tiled plane at `dev_ptr`."""
    buf = np.empty((bars_padded, 32), dtype=np.float64)
    off = block * bars_padded * 32 * 8
    rc = _cudart().cudaMemcpy(buf.ctypes.data, dev_ptr + off, buf.nbytes, 2)      # cudaMemcpyDeviceToHost
    assert rc == 0, "cudaMemcpy failed: %d" % rc
    return np.ascontiguousarray(buf[:n_bars].T)


def read_validity_rows(dev_ptr: int, s0: int, ns: int, validity_pitch: int, n_bars: int) -> np.ndarray:
    """-> bool [ns, n_bars] from the row-major Arrow bitmaps [symbol][validity_pitch bytes] at `dev_ptr`."""
    buf = np.empty((ns, validity_pitch), dtype=np.uint8)
    rc = _cudart().cudaMemcpy(buf.ctypes.data, dev_ptr + s0 * validity_pitch, buf.nbytes, 2)
    assert rc == 0, "cudaMemcpy failed: %d" % rc
    return np.unpackbits(buf, axis=1, bitorder="little")[:, :n_bars].astype(bool)

"""The C-ABI library loads on a machine without a GPU, exports every symbol include/pqb200.h
declares, and fails loudly (PQB_ERR_NO_DEVICE) instead of falling back when there is no device."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from polars_quant_b200 import _native
    _native.build()
    return _native.lib()


def declared_functions():
    text = (ROOT / "include" / "pqb200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"PQB_API\s+[\w\s\*]+?\b(pqb_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("pqb_engine_create", "pqb_panel_create", "pqb_panel_set_column", "pqb_suite_run",
                 "pqb_suite_run_host", "pqb_panel_get_output", "pqb_sma", "pqb_ema", "pqb_tema", "pqb_trima",
                 "pqb_ma", "pqb_bbands", "pqb_macd", "pqb_rsi", "pqb_trange", "pqb_atr", "pqb_natr", "pqb_obv",
                 "pqb_ad", "pqb_stoch", "pqb_kdj", "pqb_willr", "pqb_midprice", "pqb_midpoint", "pqb_adosc", "pqb_mom",
                 "pqb_roc", "pqb_cmo", "pqb_mfi", "pqb_cci", "pqb_last_error"):
        assert must in names
    assert len(names) >= 40


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert lib.pqb_abi_version() == 7
    assert lib.pqb_device_count() == 0
    h = C.c_void_p()
    rc = lib.pqb_engine_create(0, C.byref(h))
    assert rc == -1 and not h.value                      # PQB_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.pqb_last_error()


def test_default_params_are_the_reference_python_defaults(lib):
    from polars_quant_b200 import _native as N
    p = N.default_params()
    assert (p.sma_period, p.ema_period, p.tema_period, p.trima_period) == (30, 30, 30, 30)     # overlap.py
    assert (p.bbands_period, p.bbands_nbdevup, p.bbands_nbdevdn) == (20, 2.0, 2.0)
    assert (p.macd_fast, p.macd_slow, p.macd_signal, p.rsi_period) == (12, 26, 9, 14)           # momentum.py
    assert (p.atr_period, p.natr_period, p.willr_period, p.midprice_period) == (14, 14, 14, 14)
    assert (p.kdj_fastk, p.kdj_slowk, p.kdj_slowd) == (9, 3, 3)
    assert p.indicators == N.IND_ALL
    assert (p.midpoint_period, p.adosc_fast, p.adosc_slow, p.mom_period, p.roc_period) == (14, 3, 10, 10, 10)
    assert (p.cmo_period, p.mfi_period, p.cci_period) == (14, 14, 14)


def test_indicator_bits_of_the_python_mirror_match_the_header():
    """PQB_IND_* in include/pqb200.h against _native.IND / IND_EXTRA / IND_ALL / IND_FASTK (a Rust host reads the header)."""
    import re
    from polars_quant_b200 import _native as N
    text = (ROOT / "include" / "pqb200.h").read_text()
    bits = {m.group(1).lower(): 1 << int(m.group(2)) for m in re.finditer(r"PQB_IND_([A-Z]+) = 1u << (\d+)", text)}
    for name, bit in {**N.IND, **N.IND_EXTRA}.items():
        assert bits[name] == bit, name
    assert bits["fastk"] == N.IND_FASTK
    assert N.IND_ALL == sum(N.IND.values()) == (1 << 15) - 1
    assert not (N.IND_FASTK & (N.IND_ALL | sum(N.IND_EXTRA.values())))


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    for f in (ROOT / "polars_quant_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh", ".h") and f.is_file():
            assert "oracle" not in f.read_text().replace("the oracle", "").replace("oracle's", ""), f


def test_split_required_warmup_rules(lib):
    import ctypes as C
    from polars_quant_b200 import _native as N
    w = lambda **kw: lib.pqb_split_required_warmup(C.byref(N.default_params(**kw)))
    assert w() == -1                                                    # OBV / AD never forget
    assert w(indicators=N.IND["obv"]) == -1 and w(indicators=N.IND_EXTRA["midpoint"]) == -1
    assert w(indicators=N.IND["sma"]) == -1 and w(indicators=N.IND["bbands"]) == -1 and w(indicators=N.IND["kdj"]) == -1
    assert w(indicators=N.IND["willr"], willr_period=14) == 32          # a window: its own length, rounded to 32
    e = w(indicators=N.IND["ema"], ema_period=5000)
    assert e % 32 == 0 and 5000 + 14 * 5001 <= e < 5000 + 14 * 5001 + 32
    a = 2.0 / 5001.0
    assert (1.0 - a) ** (e - 5000) < 1e-12                              # the seed error is forgotten below the tolerance
    assert w(indicators=N.IND["rsi"], rsi_period=14) >= 29 * 14
    assert (1.0 - 1.0 / 14) ** (29 * 14 - 14) < 1e-12
    # groups added after the time-split code must be named explicitly to be splittable: running sums
    # (ULTOSC, CMO, MFI, CCI) are refused, the Wilder / EMA cascades ask for their own warm-up
    assert w(indicators=N.IND_EXTRA["ultosc"]) == -1 and w(indicators=N.IND_EXTRA["cmo"]) == -1
    assert w(indicators=N.IND_EXTRA["trix"], trix_period=30) >= 3 * (30 + 14 * 31)
    assert w(indicators=N.IND_EXTRA["dm"], dm_period=14) >= 2 * 29 * 14
    assert w(indicators=N.IND_EXTRA["aroon"], aroon_period=40) == 64
    assert w(indicators=1 << 27) == -1                                  # an unknown group bit is never splittable

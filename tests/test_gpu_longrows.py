"""Long rows parallel along time (BASELINE config 3): K EMAs + MACD in one pass against the serial oracle, within the
north-star tolerance (rel 1e-10 / abs 1e-12), validity exact, the first time tile bit-exact."""
import numpy as np
import pytest

import devread
import synth
import tolerances as T
from oracle import pqo

pytestmark = pytest.mark.gpu
REL, ABS = 1e-10, 1e-12


def _close_enough(name, got, ref):
    (gv, gok), (rv, rok) = got, ref
    assert np.array_equal(gok, rok), f"{name}: validity differs at {np.argwhere(gok != rok)[:4].ravel().tolist()}"
    d = np.abs(gv[rok] - rv[rok])
    lim = ABS + REL * np.abs(rv[rok])
    assert (d <= lim).all(), f"{name}: max err {d.max():.3e} at {int(np.argmax(d - lim))}, allowed {lim[np.argmax(d - lim)]:.3e}"
    return float(T.same_bits(gv[rok], rv[rok]).mean())


def test_ema_set_and_macd_against_the_serial_oracle():
    from polars_quant_b200.longrows import LongPanel
    S, NB = 70, 30_000
    d = synth.ohlcv(S, NB, seed=3, sigma=0.0005)
    for tile, periods, macd in ((256, (12, 26, 200, 5000), (12, 26, 9)), (1024, (5, 300), (3, 10, 16)), (512, (30,), None),
                                (64, (), (12, 26, 9))):
        lp = LongPanel(S, NB, ema_periods=periods, macd=macd, tile_bars=tile)
        lp.panel.set_fields(close=d["close"])
        res = lp.compute()
        for s in range(0, S, 7):
            for p in periods:
                got = (res["ema_%d" % p][0][s], res["ema_%d" % p][1][s])
                ref = pqo.ema(d["close"][s], p)
                _close_enough(f"ema({p}) tile {tile}", got, ref)
                first = max(tile, (p - 1) // tile * tile + tile)          # through the tile that holds the seed bar: the reference itself
                ok = ref[1][:first]
                assert T.same_bits(got[0][:first][ok], ref[0][:first][ok]).all(), (p, tile)
            if macd:
                for name, ref in zip(("macd", "macd_signal", "macd_hist"), pqo.macd(d["close"][s], *macd)):
                    _close_enough(f"{name}{macd} tile {tile}", (res[name][0][s], res[name][1][s]), ref)
                    ok = ref[1][:tile]
                    assert T.same_bits(res[name][0][s][:tile][ok], ref[0][:tile][ok]).all(), name
        lp.close()


def test_short_rows_and_guards():
    from polars_quant_b200.longrows import LongPanel
    from polars_quant_b200 import _native as N
    d = synth.ohlcv(3, 100, seed=4)
    lp = LongPanel(3, 100, ema_periods=(30, 200), macd=(12, 26, 9), tile_bars=64)
    lp.panel.set_fields(close=d["close"])
    res = lp.compute()
    assert not res["ema_200"][1].any()                                     # n < p: all null (overlap.rs:663)
    for s in range(3):
        for name, ref in (("ema_30", pqo.ema(d["close"][s], 30)),) + tuple(zip(("macd", "macd_signal", "macd_hist"), pqo.macd(d["close"][s], 12, 26, 9))):
            _close_enough(name, (res[name][0][s], res[name][1][s]), ref)
    lp.close()
    with pytest.raises(N.PqbError):
        LongPanel(3, 100, ema_periods=(), macd=None)
    lp = LongPanel(3, 5000, ema_periods=(10,), macd=(12, 2000, 9), tile_bars=256)
    lp.panel.set_fields(close=np.ones((3, 5000)))
    lp.panel.upload()
    with pytest.raises(N.PqbError, match="first time tile"):
        lp.run()
    lp.close()


def test_config3_full_shape_against_the_oracle():
    """500 symbols x 1,000,000 bars, EMA(12, 26, 200, 5000) + MACD(12, 26, 9), device-resident synthetic minute bars: two
    symbol blocks are read back from HBM and compared with the serial oracle over the whole million bars."""
    import polars_quant_b200 as pq
    from polars_quant_b200.longrows import LongPanel
    S, NB = 500, 1_000_000
    lib = pq._native.lib()
    lp = LongPanel(S, NB, ema_periods=(12, 26, 200, 5000), macd=(12, 26, 9), host_staging=False)
    lp.fill_synthetic(seed=3, sigma=0.0005)
    lp.run()
    lp.panel.sync()
    nb, bp = lp.panel.tiled_shape()
    h = lp.panel._h
    names = lp.names()
    for b in (0, nb - 1):
        ns = min(32, S - b * 32)
        close = devread.read_block(lib.pqb_panel_device_field(h, 0), b, bp, NB)[:ns]
        outs = {k: devread.read_block(lib.pqb_panel_device_output(h, k), b, bp, NB)[:ns] for k in names}
        oks = {k: devread.read_validity_rows(lib.pqb_panel_device_validity(h, k), b * 32, ns, lp.panel.validity_pitch, NB) for k in names}
        for s in (0, ns // 2, ns - 1):
            refs = {0: pqo.ema(close[s], 12), 1: pqo.ema(close[s], 26), 2: pqo.ema(close[s], 200), 3: pqo.ema(close[s], 5000)}
            refs.update(dict(zip((7, 8, 9), pqo.macd(close[s], 12, 26, 9))))
            for k, ref in refs.items():
                _close_enough(f"block {b} symbol {s} {names[k]}", (outs[k][s], oks[k][s]), ref)
    lp.close()

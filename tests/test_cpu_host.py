"""CPU-side tests of the product: the C-ABI library loads and exports every symbol the header
declares, the host-side switch logic reproduces the reference traces, and the product never
routes through the oracle or a CPU fallback."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def header_functions():
    text = open(os.path.join(ROOT, "include", "onda_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(onda_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from onda_b200 import _native as nat
    names = header_functions()
    assert len(names) >= 15
    lib = ctypes.CDLL(nat.lib_path())
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/onda_b200.h but not exported"
    assert sorted(nat.exported_symbols()) == names      # the ctypes binding covers the whole header
    loaded = nat.load()
    assert loaded.onda_abi_version() == 1
    # pure host helpers may be called without a GPU
    assert loaded.onda_sums_floats(19, 256) == 2 * 19 * 256 + 19 + 8
    assert loaded.onda_table_floats(19, 256) > 19 * 256


def test_no_cpu_fallback_and_no_oracle_in_product():
    from onda_b200 import prototype_handler
    pkg = os.path.join(ROOT, "onda_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no oracle", ""), f"{fn} mentions the oracle"
    h = prototype_handler(distance_metric="mahalanobis")
    h.prototypes = torch.zeros(19, 8)
    h.squared_mean = torch.ones(19, 8)
    h.counter = torch.ones(19)
    with pytest.raises(RuntimeError, match="CUDA"):
        h.pseudo_labels(torch.zeros(1, 8, 2, 2), torch.ones(1, 19, 2, 2))
    with pytest.raises(RuntimeError, match="CUDA"):
        h.ma(torch.zeros(1, 8, 2, 2), torch.ones(1, 19, 2, 2))
    with pytest.raises(ValueError):
        prototype_handler(distance_metric="cosine")
    with pytest.raises(AttributeError):
        h.pseudo_labels(torch.zeros(1, 8, 2, 2), None)


def test_transform_and_onehot_helpers():
    from onda_b200 import prototype_handler
    h = prototype_handler()
    x = torch.arange(2 * 3 * 2 * 2, dtype=torch.float32).reshape(2, 3, 2, 2)
    rows = h.transform(x)
    assert rows.shape == (8, 3) and torch.equal(rows[1], x[0, :, 0, 1])
    assert h.transform(rows) is rows
    hot = h.onehot(torch.tensor([[1.0, 3.0, 3.0, 2.0]]))
    assert hot.tolist() == [[0.0, 1.0, 0.0, 0.0]]       # first maximal index wins


@pytest.mark.parametrize("tag", ["hamming", "median", "mean"])
def test_switch_logic_matches_reference_traces(tag):
    from onda_b200 import Monitor, HybridSelect, DevSelect, static_share
    z = np.load(os.path.join(GOLDEN, f"monitor_trace_{tag}.npz"))
    mon = Monitor(int(z["limit"]), float(z["exp_const"]), str(z["dev_func"]))
    sel = HybridSelect(HybridSelect.static, tuple(z["gray"]), float(z["dev_thresh"]))
    vsel = DevSelect(DevSelect.static, float(z["vthresh"]))
    cur, cur_dev, dev, med, ema, vcur, pct = [], [], [], [], [], [], []
    for v in z["conf"]:
        mon.add({"prior static": float(v)})
        d = mon.dev_avg("prior static")
        sel.evaluate(mon.avg("prior static"), d)
        vsel.evaluate(d)
        cur.append(sel.current); cur_dev.append(sel.current_dev); dev.append(d)
        med.append(mon.avg("prior static")); ema.append(mon.exp("prior static")); vcur.append(vsel.current)
        pct.append(static_share(mon.avg("prior static"), True))
    assert np.array_equal(cur, z["ref_current"]) and np.array_equal(cur_dev, z["ref_current_dev"])
    assert np.array_equal(vcur, z["ref_vcurrent"])
    assert np.array_equal(np.array(dev, dtype=np.float64), z["ref_dev"])
    assert np.array_equal(np.array(med), z["ref_median"]) and np.array_equal(np.array(ema), z["ref_exp"])
    assert np.array_equal(np.array(pct), z["ref_pct"])
    empty = Monitor(5)
    assert empty.avg("nope") == z["ref_missing_avg"] and empty.exp("nope") == z["ref_missing_exp"]
    assert empty.dev_avg("nope") == z["ref_missing_dev"]
    # frozen selectors and monitors ignore updates
    sel.eval(); before = sel.current; sel.evaluate(0.0, 1.0); assert sel.current == before
    mon.eval(); n = len(mon.current_dict["prior static"]); mon.add({"prior static": 0.5})
    assert len(mon.current_dict["prior static"]) == n
    assert static_share(0.9, False, 0.85) == 1 and static_share(0.8, False, 0.85) == 0

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


def test_load_reads_three_tuple_and_legacy_two_tuple(tmp_path):
    """save/load keep the reference's 3-tuple pickle (prototype_handler.py:37-47); the legacy 2-tuple shipped as
    prototypes.pickle (prototypes, counter) loads too, with squared_mean left uninitialised."""
    import pickle
    from onda_b200 import prototype_handler
    P, S, c = torch.randn(19, 8), torch.rand(19, 8) + 1, torch.arange(19.0)
    h = prototype_handler()
    h.prototypes, h.squared_mean, h.counter = P, S, c
    loc = str(tmp_path / "state.pickle")
    h.save(loc)
    assert [torch.equal(a, b) for a, b in zip(pickle.load(open(loc, "rb")), (P, S, c))] == [True] * 3
    h2 = prototype_handler()
    assert h2.load(loc) and torch.equal(h2.prototypes, P) and torch.equal(h2.squared_mean, S) and torch.equal(h2.counter, c)
    legacy = str(tmp_path / "legacy.pickle")
    pickle.dump((P, c), open(legacy, "wb"))
    h3 = prototype_handler()
    assert h3.load(legacy) and torch.equal(h3.prototypes, P) and torch.equal(h3.counter, c)
    assert isinstance(h3.squared_mean, int) and h3.squared_mean == 0
    assert not prototype_handler().load(str(tmp_path / "absent.pickle"))


def test_oracle_step_log_stats_small_case():
    """The restated log reductions on a hand-checkable case (prototypes.py:341-352)."""
    from oracle import proto_oracle as po
    labels = torch.tensor([[0], [1], [255], [2]])
    out = torch.zeros(1, 3, 2, 2)
    out[0, 0, 0, 0] = 1      # pixel 0 -> class 0 (agrees)
    out[0, 2, 0, 1] = 1      # pixel 1 -> class 2 (label 1: disagrees)
    out[0, 1, 1, 0] = 1      # pixel 2 -> class 1 (label 255: never agrees)
    out[0, 2, 1, 1] = 1      # pixel 3 -> class 2 (agrees)
    got = po.step_log_stats(labels, out, torch.tensor([[1.0, 2.0], [3.0, 4.0]]))
    assert got == {"pseudolabel_pixel_num": 3.0, "output & prototype agreement": 0.5, "mean_prototype_intensity_values": 7.5}


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) needs no GPU and prints one JSON line
    with the keys of the bench contract."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "px/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "px/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["vs_baseline"] is None


def test_oracle_update_ema_small_case():
    """The restated weight EMA (prototypes.py:407-416): parameters blended, buffers copied."""
    from oracle import proto_oracle as po
    new_p, new_b = po.update_ema([torch.tensor([2.0, 4.0])], [torch.tensor([0.0, 8.0])],
                                 [torch.tensor([7])], [torch.tensor([1])], 0.75)
    assert new_p[0].tolist() == [0.5, 7.0] and new_b[0].tolist() == [7]


def test_oracle_eval_confusion_small_case():
    """The restated evaluation counters (adaptation_model.py:143-160, func.py:77-79) on a hand-checkable case."""
    from oracle import proto_oracle as po
    pred = torch.zeros(1, 3, 2, 2)
    pred[0, 1, 0, 0] = 5     # top-left -> class 1
    pred[0, 2, 0, 1] = 5     # top-right -> class 2
    pred[0, 0, 1, 0] = 5     # bottom-left -> class 0
    pred[0, 2, 1, 1] = 5     # bottom-right -> class 2
    labels = torch.tensor([[[1, 2], [255, 0]]])
    _, p, hist = po.eval_confusion(pred, labels, 3, (2, 2))
    assert p.tolist() == [[[1, 2], [0, 2]]]
    assert hist.tolist() == [[0, 0, 1], [0, 1, 0], [0, 0, 1]]     # (label 0 -> pred 2), (1 -> 1), (2 -> 2); 255 ignored

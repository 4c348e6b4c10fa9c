"""scripts/timeline_decode.py: the trace analysis (pure host code) on a synthetic kernel trace."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load():
    spec = importlib.util.spec_from_file_location("timeline_decode", os.path.join(ROOT, "scripts", "timeline_decode.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_short_name():
    m = _load()
    assert m.short_name("void sv::mega::gemv_ring_kernel<true, 1, false>(sv::mega::RingGemvArgs)") == "gemv_ring_kernel<true, 1, false>"
    assert m.short_name("sv::attention_decode_cluster_kernel(const bf16*)") == "attention_decode_cluster_kernel"


def test_analyse_steps_gaps_and_overlap():
    m = _load()
    ev, t = [], 100.0
    for step in range(5):
        # gemv 10us, then attention starts 2us BEFORE gemv ends (overlap), then select after a 3us gap
        ev.append({"name": "void sv::gemv_ring_kernel<true>(int)", "ts": t, "dur": 10.0})
        ev.append({"name": "sv::attention_kernel(int)", "ts": t + 8.0, "dur": 5.0})
        ev.append({"name": "sv::select_fused_kernel(int)", "ts": t + 16.0, "dur": 1.0})
        t += 20.0
    out = m.analyse(ev, skip_steps=1)
    assert out["steps"] == 4 and out["launches_per_step"] == 3
    assert abs(out["step_us"] - 17.0) < 1e-9 and abs(out["busy_us"] - 14.0) < 1e-9 and abs(out["idle_us"] - 3.0) < 1e-9
    rows = {r["kernel"]: r for r in out["kernels"]}
    assert abs(rows["attention_kernel"]["overlap_us_per_step"] - 2.0) < 1e-9
    assert abs(rows["select_fused_kernel"]["gap_us_per_step"] - 3.0) < 1e-9
    assert abs(out["sum_kernel_us"] - 16.0) < 1e-9

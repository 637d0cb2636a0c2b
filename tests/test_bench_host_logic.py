"""Host logic of bench.py that runs without a GPU: the merge of the per-rank extras (maximum over ranks of every
timing, throughput of all ranks over that time, rank-0 samples passed through) and the constants the roofline entries
are built from."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_merge_extras_takes_the_slowest_rank_and_counts_every_ranks_points():
    b = _bench()
    n, world = 1_000_000, 2
    r0 = {"linearize_100M": {"kannala_brandt/pixel": {"ms": 1.0, "_bytes_per_point": 40, "sustained_ms": 1.25, "_r0_sm_mhz": 1700.0}},
          "project_unproject_100M_f64": {"ucm": {"project_ms": 0.5, "project_bytes_per_point": 41, "unproject_ms": 0.4, "unproject_bytes_per_point": 41}}}
    r1 = {"linearize_100M": {"kannala_brandt/pixel": {"ms": 2.0, "_bytes_per_point": 40, "sustained_ms": 1.0}},
          "project_unproject_100M_f64": {"ucm": {"project_ms": 0.25, "project_bytes_per_point": 41, "unproject_ms": 0.8, "unproject_bytes_per_point": 41}}}
    m = b.merge_extras([r0, r1], n, world)
    kb = m["linearize_100M"]["kannala_brandt/pixel"]
    assert kb["ms"] == 2.0 and kb["sustained_ms"] == 1.25          # maximum over ranks
    assert kb["sm_mhz"] == 1700.0                                  # sampled on rank 0 only
    assert kb["gpts_s"] == n * world / 2.0 / 1e6                   # all ranks' points over the slowest rank's time
    assert kb["gb_s"] == n * 40 / 2.0 / 1e6                        # per GPU
    assert kb["gb_s_all_gpus"] == world * kb["gb_s"]
    u = m["project_unproject_100M_f64"]["ucm"]
    assert u["project_ms"] == 0.5 and u["unproject_ms"] == 0.8
    assert u["unproject_gb_s"] == n * 41 / 0.8 / 1e6


def test_fp64_issue_constants():
    b = _bench()
    assert b.FP64_LANES_PER_CLOCK == 148 * 64
    assert set(b.FP64_INSTR_PER_POINT) == {"double_sphere/pixel", "kannala_brandt/pixel", "rad_tan/pixel", "fov/pixel"}
    # 102 FP64 instructions per point at 100 G points/s and 1.5 GHz is 72 % of the pipe
    frac = 102 * 100e9 / (b.FP64_LANES_PER_CLOCK * 1.5e9)
    assert 0.71 < frac < 0.73
    assert b.BYTES_PER_POINT == 40

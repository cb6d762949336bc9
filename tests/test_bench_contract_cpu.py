"""The bench line contract, as far as it can be checked without a GPU: the reference arm runs here (the
unmodified reference from baseline/_ref when installed, the C port otherwise) and must print ONE JSON line with
the agreed keys; both arms print the same `config` object."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, GMZ_BENCH_MAX_PROCS="2")           # keep the CPU suite short: two reference processes
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-topology-seconds", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "mcts_sims_per_sec" and d["unit"] == "sims/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    cb = d["cpu_baseline"]
    sys.path.insert(0, ROOT)
    from baseline import ref_runner
    if ref_runner.available():          # the real reference is the headline, the port the second figure
        assert cb["kind"] == "reference" and cb["port"]["kind"] == "port" and cb["port"]["value"] > cb["value"]
        assert "config1" in cb
    else:
        assert cb["kind"] == "port"
    assert cb["cores"] >= 1 and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    import bench
    assert d["config"] == bench.bench_config(4096)            # byte-identical to what the GPU arm prints


def test_both_arms_share_the_config_object():
    sys.path.insert(0, ROOT)
    import bench
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": bench_config(') == 2          # our arm and the reference arm
    assert "15x15" in bench.WORKLOAD and "400 sims" in bench.WORKLOAD and "4096" in bench.WORKLOAD
    assert bench.own_bytes_per_sim(3.8) < bench.algorithmic_bytes_per_sim(3.8)


def test_reference_install_is_unmodified():
    """baseline/_ref (when present) holds byte-identical copies of the reference modules (SHA-256 manifest)."""
    sys.path.insert(0, ROOT)
    from baseline import install_reference, ref_runner
    if ref_runner.available():
        assert install_reference.verify()


def test_reference_in_process_network_runner():
    """one reference process with the reference's own network answering its queue
    protocol in-process -- a few moves at a tiny size, in both search modes (on the CPU here, on the B200 in the bench)."""
    sys.path.insert(0, ROOT)
    import pytest
    from baseline import ref_runner
    if not ref_runner.available():
        pytest.skip("baseline/_ref not installed")
    for mode in ("AlphaZero", "MuZero"):
        r = ref_runner.inprocess_net(seconds=1.0, N=6, S=16, K=4, mode=mode)
        assert r["moves"] >= 1 and r["sims_per_sec"] > 0 and r["evaluator_calls"] >= r["moves"] * (16 if mode == "AlphaZero" else 2)
        assert mode in r["what"]


def test_traffic_stamp_is_current_and_ignores_comments():
    """`roofline.traffic` is an ncu measurement: profiles/traffic.json names the play-kernel sources it was taken on
    (comments and blank lines do not count) and must match the sources in the tree."""
    sys.path.insert(0, ROOT)
    import bench
    tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert tj["kernel_source_sha"] == bench.kernel_source_sha()
    src = {f: open(os.path.join(ROOT, "datou_gomoku_muzero_b200", "csrc", f)).read() for f in bench.PLAY_KERNEL_SOURCES}
    commented = dict(src, **{"gmz_play.cuh": "// a remark\n\n" + src["gmz_play.cuh"] + "\n/* another\n one */\n"})
    changed = dict(src, **{"gmz_play.cuh": src["gmz_play.cuh"].replace("GMZ_PLAY_MIN_CTAS 28", "GMZ_PLAY_MIN_CTAS 27")})
    assert bench.kernel_source_sha(commented.__getitem__) == bench.kernel_source_sha()
    assert bench.kernel_source_sha(changed.__getitem__) != bench.kernel_source_sha()

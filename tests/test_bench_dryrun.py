"""CPU tests of bench.py's GPU arm CONTROL FLOW (tests/bench_dryrun_driver.py replaces the device with stand-ins): the one JSON
line, what happens when a batch fails on one rank only (nobody may be left waiting in a collective), and the deadline of the
extra configurations.  No number printed by these runs means anything."""
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "bench_dryrun_driver.py")
ARGS = ["--frames", "4", "--steps", "3", "--warmup", "3", "--no-cpu", "--parity-frames", "0"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, extras, env_extra, timeout=240):
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, **env_extra)
        if world > 1:
            env.update(RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        else:
            for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
                env.pop(k, None)
        procs.append(subprocess.Popen([sys.executable, DRIVER, "--gpus", str(world), "--extras", extras] + ARGS, cwd=ROOT, env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        try:
            o, e = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise AssertionError("a rank was left waiting (no exit after %d s)" % timeout)
        outs.append((p.returncode, [l for l in o.splitlines() if l.strip()], e))
    return outs


def test_one_json_line_with_the_contract_keys():
    (rc, lines, err), = _run(1, "", {})
    assert rc == 0, err[-2000:]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "e2e", "gpu_launches", "roofline", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["e2e"]["h2d_bytes_per_step"] > 0 and "workload" in d["config"]
    assert d["pipeline"]["one_batch_at_a_time"]["e2e_jpeg_value"] is not None


def test_a_failing_secondary_pass_is_reported_as_null():
    (rc, lines, err), = _run(1, "", {"DRYRUN_FAIL": "jpeg:1:0"})
    assert rc == 0, err[-2000:]
    d = json.loads(lines[-1])
    assert d["value"] > 0 and d["e2e"]["value"] > 0
    assert d["pipeline"]["one_batch_at_a_time"]["e2e_jpeg_value"] is None and d["pipeline"]["one_batch_at_a_time"]["value"] is not None


def test_a_failing_headline_pass_fails_loudly():
    (rc, lines, err), = _run(1, "", {"DRYRUN_FAIL": "dev:2:0"})
    assert rc != 0 and not lines and "bench step failed" in err


def test_two_ranks_one_fails_nobody_hangs():
    outs = _run(2, "", {"DRYRUN_FAIL": "jpeg:1:1"})        # rank 1 fails in the one-batch-at-a-time JPEG pass
    assert [o[0] for o in outs] == [0, 0], outs[0][2][-1500:] + outs[1][2][-1500:]
    assert len(outs[0][1]) == 1 and not outs[1][1]          # rank 0 alone prints
    d = json.loads(outs[0][1][0])
    assert d["n_gpus"] == 2 and d["pipeline"]["one_batch_at_a_time"]["e2e_jpeg_value"] is None and d["value"] > 0
    assert "on another rank" in outs[0][2]


def test_two_ranks_headline_failure_ends_both():
    outs = _run(2, "", {"DRYRUN_FAIL": "jpeg:2:1"})        # the e2e headline pass (2 contexts) fails on rank 1
    assert all(o[0] != 0 for o in outs) and not outs[0][1]


def test_extras_deadline_keeps_the_headline():
    outs = _run(2, "c5", {"DRYRUN_C5": "hang"})            # rank 1 dies inside c5, rank 0 waits for it forever
    assert outs[0][0] == 0 and len(outs[0][1]) == 1
    d = json.loads(outs[0][1][0])
    assert d["value"] > 0 and "error" in d["extra_configs"]["c5"]

"""The bench.py JSON contract, checked on the CPU-runnable arm (`--impl
reference`): one JSON line, the keys the driver reads, and the rule that ranks
other than 0 stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env=None, *extra):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "16",
                        "--cpu-seconds", "1", "--steps", "2", "--warmup", "1", *extra],
                       capture_output=True, text=True, env=e, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip()


def test_reference_arm_json_line():
    out = run()
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "higher_is_better", "scaling",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["unit"] == "DoFs/s" and d["dtype"] == "f64" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_are_silent():
    assert run({"RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2") == ""


def test_counted_work_follows_from_the_committed_sass_listing():
    """bench.py's roofline uses profiles/kernel_counts.json; the numbers there must be what
    tools/sass_loops.py counts in the committed listing of the headline kernel (one pipeline unit of
    8 cells = the main loop once + the rolled zeta loop four times)."""
    import re
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_loops
    ins = sass_loops.parse(os.path.join(ROOT, "profiles", "r02_action_cg3.sass"))
    loops = []
    for a, t, _ in ins:
        m = re.match(r"BRA(?!\.DIV)\S*\s+(?:\S+,\s*)?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            loops.append((int(m.group(1), 16), a))
    outer = max(loops, key=lambda l: l[1] - l[0])
    inner = max((l for l in loops if outer[0] <= l[0] and l[1] < outer[1]), key=lambda l: l[1] - l[0])
    count = {}
    for a, t, _ in ins:
        if outer[0] <= a <= outer[1]:
            op = t.split()[0].split(".")[0]
            count[op] = count.get(op, 0) + (4 if inner[0] <= a <= inner[1] else 1)
    k = json.load(open(os.path.join(ROOT, "profiles", "kernel_counts.json")))["helmholtz_action_kernel<4,false,true,3>"]
    assert (count["DFMA"], count["DMUL"], count["DADD"], count["MUFU"]) == (k["dfma"], k["dmul"], k["dadd"], k["mufu_rcp64h"])
    fp64 = count["DFMA"] + count["DMUL"] + count["DADD"] + count["MUFU"]
    assert fp64 == k["fp64_instr_per_warp_unit"] and sum(count.values()) == k["instr_per_warp_unit"]
    assert k["fp64_instr_per_cell"] == fp64 * 32 // k["cells_per_warp_unit"]
    assert k["flop_per_cell"] == (2 * count["DFMA"] + count["DMUL"] + count["DADD"]) * 32 // k["cells_per_warp_unit"]

"""bench.py's driver-facing contract, the part that runs without a GPU: the
reference arm prints exactly one JSON line on stdout with the agreed keys."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    proc = subprocess.run(
        [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "2",
         "--warmup", "1", "--docs", "60000", "--vocab", "50000", "--ref-real-docs", "20000"],
        capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] >= 1 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("C2:") and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["vs_baseline"] is None and d["gpu_launches"] == 0
    cr = d.get("compiled_reference")
    if cr and "unavailable" not in cr:
        assert cr["reference_queries_per_s_1core"] > 0 and cr["port_queries_per_s_1core"] > 0


def test_other_ranks_of_the_reference_arm_do_no_work():
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    proc = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2"],
                          capture_output=True, text=True, timeout=120, env=env)
    assert proc.returncode == 0 and proc.stdout.strip() == ""


def test_bench_query_strings_compile_to_the_programs_the_device_leg_uses():
    """bench.py scores resident batches from (tokens, program) lists and runs the
    e2e leg from query strings: for every C2 / C3 template the host parser must
    turn the string into exactly the token order and postfix program the
    generator wrote (ref grammar.y precedence NOT > AND > OR, query.c:89-103
    right-to-left token order)."""
    import numpy as np
    import bench
    from nxsearch_b200 import tools

    corpus = tools.Corpus.generate(2_000, 3_000)
    qt = corpus.query_terms(6 * 64)
    for shape in ("or", "bool"):
        for item in bench.make_queries(qt, 64, shape):
            toks, prog, _ = item
            compiled = tools.query_compile(bench.query_string(corpus, item))
            assert compiled is not None
            leaves, cprog = compiled
            assert [corpus.term(t) for t in toks] == leaves, bench.query_string(corpus, item)
            assert list(prog) == list(cprog), bench.query_string(corpus, item)

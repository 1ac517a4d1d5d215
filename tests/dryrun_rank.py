"""One rank of tests/test_bench_dryrun.py::test_two_rank_glue_over_gloo: bench.py's own arm on the CPU stand-in engine."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import bench  # noqa: E402
import fake_engine  # noqa: E402

fake_engine.install(bench)
if os.environ.get("DRYRUN_BIG"):
    bench.BIG_CORPUS_BYTES = 1e4
bench.main()

"""The source-level drop-in check on the GPU: tests/c/dropin_usearch_h.c is compiled against the REFERENCE's own usearch.h by
__graft_entry__.build() (only possible where /root/reference exists) and travels to the GPU box as a binary.  It plays
build.c's and scan.c's call sequence (usearch_init / reserve / add / search_ef / save_buffer / update_header / count / ...)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "c", "_bin", "dropin_usearch_h")


def test_caller_built_against_the_reference_header_runs_on_the_gpu(eng):
    if not os.path.exists(EXE):
        pytest.skip("tests/c/_bin/dropin_usearch_h not prebuilt (build() compiles it where the reference header exists)")
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "drop-in ok: nearest key 102" in r.stdout, (r.returncode, r.stdout, r.stderr)

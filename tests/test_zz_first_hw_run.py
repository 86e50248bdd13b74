"""First hardware run of the kernels that were written while no GPU was available (attention v4 / v5, post-processing).

They are opt-in (DSHEG_ATTN=..., DSHEG_RUN_UNVALIDATED=1) and validated on the CPU emulator (tests/test_emu_kernels.py);
this test (opt-in: DSHEG_FIRST_HW_RUN=1) runs scripts/first_hw_run.py in a SUBPROCESS with a timeout -- a faulting or hanging kernel cannot poison the CUDA
context of the parity suite, and the file sorts last so everything else has already run -- prints its report, and is
an expected-failure on error: the shipped default path does not depend on any of these kernels, so a failure here is
information for the next round, not a broken build.  Once a run is green the gates come off (see profiles/r01/NOTES_next_round.md)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get("DSHEG_FIRST_HW_RUN") != "1",
                    reason="opt-in (DSHEG_FIRST_HW_RUN=1): never-before-executed kernels are not run inside an unattended GPU suite -- "
                           "a hung cluster barrier would cost the whole box session; scripts/gpu_round2_first.sh runs them under its own timeouts")
def test_first_hardware_run_of_unvalidated_kernels():
    try:
        # parity only (B = 3): the B = 950 bandwidth comparison belongs to scripts/gpu_round2_first.sh, not to the test suite
        r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "first_hw_run.py")], cwd=ROOT, capture_output=True,
                           text=True, timeout=1500, env=dict(os.environ, DSHEG_FIRST_RUN_BATCH="0"))
    except subprocess.TimeoutExpired as e:
        print((e.stdout or b"").decode(errors="replace")[-3000:] if isinstance(e.stdout, bytes) else (e.stdout or "")[-3000:])
        pytest.xfail("first hardware run timed out after 1500 s (see the report above)")
    print("\n[first hardware run]\n" + r.stdout[-6000:])
    if r.returncode != 0:
        print(r.stderr[-3000:])
        pytest.xfail("a never-before-executed opt-in kernel failed its first hardware run (report above); default path unaffected")

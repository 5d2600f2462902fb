"""CPU check of the warp-strip Gauss-Seidel kernel's SCHEDULE (saras_b200/csrc/gs_strip.cuh): scripts/sim/gs_strip_emul.py
restates the kernel's position / ring-slot / dependency logic lane by lane and executes it with real fp64 arithmetic under
random warp interleavings and random completion times of its asynchronous copies; the result must equal a sequential
lexicographic Gauss-Seidel sweep bit for bit.  The CUDA kernel itself is compared with the oracle on the GPU
(tests/test_gpu_parity.py::test_gs_strip_*)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts", "sim"))
import gs_strip_emul as emul  # noqa: E402


@pytest.mark.parametrize("case", emul.CASES + [(3, 40, 9, 2, 2, 2, 32), (4, 9, 21, 2, 2, 3, 4)], ids=lambda c: "x".join(map(str, c)))
def test_strip_schedule_is_lexicographic_gauss_seidel(case):
    for seed in range(4):
        assert emul.run_case(*case, seed, slots=(10 ** 9 if seed % 2 else 2), zuni=(seed % 3 != 1))


def test_emulation_catches_a_missing_dependency():
    """the check has teeth: without the -y wait the sweep must NOT come out right"""
    src = open(emul.__file__).read()
    mutated = src.replace('needY = pym + LB if (has_y and 0 <= pym < Pend) else 0', 'needY = 0')
    assert mutated != src
    ns = {}
    exec(compile(mutated.replace('if __name__ == "__main__":\n    sys.exit(main())', ''), "emul_mutated", "exec"), ns)
    bad = 0
    for seed in range(6):
        try:
            ns["run_case"](5, 9, 17, 2, 2, 2, 4, seed)
        except (AssertionError, TypeError):
            bad += 1
    assert bad >= 3

"""The schedule of the planned strip Gauss-Seidel sweep keeps the reference's update order under any
interleaving of its warps and never deadlocks -- checked with the version-tracking simulator on a small subset of the
shapes `python scripts/sim/gs_strip_sim.py` sweeps."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("gs_strip_sim", os.path.join(ROOT, "scripts", "sim", "gs_strip_sim.py"))
sim = importlib.util.module_from_spec(spec)
spec.loader.exec_module(sim)


@pytest.mark.parametrize("shape", [(5, 9, 9), (7, 10, 17), (4, 3, 5)])
@pytest.mark.parametrize("W", [1, 3])
@pytest.mark.parametrize("slots", [1, 1000])
def test_strip_schedule_orders_reads_and_terminates(shape, W, slots):
    for seed in range(2):
        sim.Sim(*shape, LB=4, W=W, NZ=2, R=1, G=3, slots=slots, seed=seed).run()


def test_ring_shallower_than_one_lap_deadlocks():
    """the design constraint the simulator found: the -x ring must hold at least ceil(S / W) steps"""
    s = sim.Sim(5, 9, 9, LB=2, W=1, NZ=1, R=-1, G=1, slots=1, seed=0)      # depth = ceil(S/W) - 1
    with pytest.raises(AssertionError, match="deadlock"):
        s.run()

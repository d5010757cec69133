"""bench.py's control flow, executed without a GPU (tools/bench_emu.py): the emulator build of the
library, host stand-ins for torch's CUDA entry points, gloo instead of NCCL.  Checks that every leg
runs, that the self-checks inside the bench pass (oracle parity, gather content, host paths against
the device path) and that the one JSON line has the contract's keys.  Nothing here is a measurement."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


WF = ["--workload", "C5", "--wf-frame", "64x32x2", "--scale", "0.0005"]      # the wavefront leg on a toy frame


@pytest.mark.parametrize("world,extra", [(1, []), (2, []), (2, ["--gather", "nccl"]), (1, WF), (2, WF)])
def test_bench_dry_run(world, extra):
    """the second case is the peer-memory gather (the default) between two real processes: the emulator then
    backs 'device memory' with POSIX shared memory and its CUDA IPC calls map the window for real; the third is
    the torch.distributed gather"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_emu.py"), str(world)] + extra,
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert f"dry run ok (world {world}" in r.stdout

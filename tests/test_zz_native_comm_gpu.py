"""vvae_comm_* on a GPU (csrc/comm.cu, ddp.NativeComm): a one-rank communicator through the C ABI.

Written after this round's GPU budget was spent, so the round-end run is its first execution: it runs LAST (file name),
in a CHILD process with a time limit, and is a non-strict xfail -- whatever NCCL does on the box, it cannot stop or
poison the parity suite.  The two-rank path has not been run at all (DESIGN.md section 5 says so)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, torch
sys.path.insert(0, %r)
from video_vae_b200 import ddp
torch.cuda.set_device(0)
comm = ddp.NativeComm(0, 1)
x = torch.arange(1 << 20, dtype=torch.float32, device="cuda")
y = x.clone()
comm.all_reduce(y)                    # sum over one rank
comm.all_reduce(y, average=True)      # mean over one rank
comm.broadcast(y, 0)
z = y.to(torch.bfloat16)
comm.all_reduce(z)
torch.cuda.synchronize()
assert torch.equal(x, y) and torch.equal(z, x.to(torch.bfloat16))
comm.close()
print("native comm ok")
""" % ROOT


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="first executed by the round-end run (see the module docstring)")
def test_native_comm_single_rank_gpu():
    r = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, timeout=240)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "native comm ok" in r.stdout

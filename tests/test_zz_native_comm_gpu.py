"""vvae_comm_* on a GPU (csrc/comm.cu, ddp.NativeComm): a one-rank communicator through the C ABI.

Runs in a CHILD process with a time limit (NCCL initialisation is the one blocking call of the library; a wedged
communicator must not take the parity suite with it).  Passed on a B200 (profiles/r02zzz_native_comm_n1_pytest.log); two
ranks: scripts/check_native_comm.py under torchrun, profiles/r02zzz_native_comm_n2.log (sum / mean / broadcast exact, the
682 MB production gradient all-reduce in 1.28 ms)."""
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
def test_native_comm_single_rank_gpu():
    r = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, timeout=240)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "native comm ok" in r.stdout

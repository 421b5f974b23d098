"""Multi-rank correctness of the displayed-frame path on whatever GPUs the box has: one process per rank under
torch.distributed.run, the ranks spread over the visible devices (sharing one GPU when there is only one - the frame
path needs no collective, only CUDA IPC), the assembled frame byte-identical to a single-rank frame."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n_ranks,width", [(2, 320), (3, 208), (2, 333), (4, 1920)])
def test_ranks_assemble_the_single_rank_frame(n_ranks, width):
    env = dict(os.environ, RT_TEST_WIDTH=str(width))
    port = 29650 + n_ranks + width % 7
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n_ranks}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(REPO, "tests", "multirank_frame_worker.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert "MULTIRANK_FRAMES_OK" in r.stdout, r.stdout[-2000:]

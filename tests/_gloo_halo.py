"""Host-side halo exchange over torch.distributed (gloo) for the CPU tests of
the partition / halo INDEX LISTS (the device path uses NCCL in halo.cu).  Test
infrastructure: never imported by the package."""
import numpy as np
import torch
import torch.distributed as dist


def exchange(neighbours, data, reverse):
    """reverse=False: owners -> ghosts (REPLACE); True: ghosts += into owners."""
    reqs, recvs = [], []
    for rank, send, recv in neighbours:
        s, r = (recv, send) if reverse else (send, recv)
        if len(s):
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(data[s])), rank))
        if len(r):
            buf = torch.empty(len(r), dtype=torch.float64)
            reqs.append(dist.irecv(buf, rank))
            recvs.append((r, buf))
    for q in reqs:
        q.wait()
    for idx, buf in recvs:
        if reverse:
            data[idx] += buf.numpy()
        else:
            data[idx] = buf.numpy()

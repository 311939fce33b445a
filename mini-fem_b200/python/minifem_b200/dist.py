"""One process per GPU: torch.distributed is the plumbing (rendezvous, the NCCL unique id,
timer reduction); the interface sum itself runs inside libminifem_b200 over NCCL.  A host
transport (gloo or nccl point-to-point through torch) exists for tests and debugging."""
import os

import numpy as np
import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT as torchrun sets them."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def broadcast_bytes(payload, length, src=0):
    """Every rank gets rank `src`'s `length` bytes."""
    if not dist.is_initialized():
        return payload
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(length, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def comm_init(ctx):
    """mfb_comm_unique_id on rank 0, broadcast, mfb_ctx_comm_init everywhere."""
    from . import comm_unique_id, COMM_ID_BYTES
    if not dist.is_initialized():
        return
    uid = comm_unique_id() if dist.get_rank() == 0 else bytes(COMM_ID_BYTES)
    uid = broadcast_bytes(uid, COMM_ID_BYTES, src=0)
    ctx.comm_init(uid)


def all_gather_bytes(payload, length):
    """Every rank's `length` bytes, in rank order."""
    if not dist.is_initialized():
        return [bytes(payload)]
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    parts = [torch.zeros(length, dtype=torch.uint8, device=dev) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, mine)
    return [bytes(t.cpu().numpy().tobytes()) for t in parts]


def p2p_connect(ctx):
    """Peer-to-peer exchange of the fused RING iteration (mfb_ctx_p2p_*): every rank publishes its card, the cards
    travel through torch.distributed, every rank maps its neighbours' windows.  All ranks switch together: if any of
    them cannot connect (no peer access, cudaIpc refused, more than 64 interfaces) everybody stays on NCCL.  Returns
    (active, reason).  MFB_HALO=nccl keeps the NCCL exchange."""
    from . import MfbError, P2P_CARD_BYTES
    if os.environ.get("MFB_HALO", "p2p").lower() == "nccl":
        return False, "MFB_HALO=nccl"
    why = ""
    try:
        card = ctx.p2p_card()
    except MfbError as e:
        card, why = bytes(P2P_CARD_BYTES), str(e)
    cards = all_gather_bytes(card, P2P_CARD_BYTES)
    ok = 0.0
    if not why:
        try:
            ctx.p2p_connect(cards)
            ok = 1.0
        except MfbError as e:
            why = str(e)
    everyone = -max_over_ranks(-ok)                     # min over ranks
    if everyone < 1.0:
        if ok:
            ctx.p2p_enable(False)
        return False, why or "another rank could not connect"
    return True, ""


def max_over_ranks(value):
    if not dist.is_initialized():
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value):
    if not dist.is_initialized():
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if dist.is_initialized():
        dist.barrier()


def exchange_host(sendbuf, intfIndex, neighborsList, dim):
    """The message pattern of MPI_halo_exchange (halo.cc:52-96) over torch.distributed:
    segment i of `sendbuf` goes to rank neighborsList[i]-1, and what that rank sends lands
    in segment i of the returned buffer."""
    recvbuf = np.zeros_like(sendbuf)
    if not dist.is_initialized():
        return recvbuf
    on_gpu = dist.get_backend() == "nccl"
    send_t = torch.from_numpy(np.ascontiguousarray(sendbuf))
    recv_t = torch.zeros_like(send_t)
    if on_gpu:
        send_t, recv_t = send_t.cuda(), recv_t.cuda()
    ops = []
    for i in range(len(intfIndex) - 1):
        lo, hi = int(intfIndex[i]) * dim, int(intfIndex[i + 1]) * dim
        peer = int(neighborsList[i]) - 1
        ops.append(dist.P2POp(dist.irecv, recv_t[lo:hi], peer))
        ops.append(dist.P2POp(dist.isend, send_t[lo:hi], peer))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    recvbuf[:] = recv_t.cpu().numpy()
    return recvbuf

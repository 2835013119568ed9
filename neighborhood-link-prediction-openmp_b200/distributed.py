"""Multi-GPU plumbing: one process per GPU, replicated CSR, source vertices partitioned by
``nlp_set_partition``; every rank produces its local top-K, ONE all-gather moves the candidates
(NCCL over NVLink on GPUs, gloo in the CPU tests), and the same on-device select that ends a
single-GPU prediction picks the global top-K.  Replaces the serial T-way heap merge of
inc/predict.hxx:431-460.  The canonical (score desc, u asc, v asc) order makes the result
independent of the number of ranks.
"""
import torch
import torch.distributed as dist


def owner_of_vertex(u, world):
    """Rank that owns source vertex ``u`` (mirror of owns_row_block in csrc/frontier.cuh):
    blocks of 32 consecutive ids are dealt round-robin."""
    return (u // 32) % world


def gather_candidates(u, v, s, group=None):
    """All-gather variable-length candidate lists.

    ``u``, ``v`` (int32) and ``s`` (float32) are this rank's local top-K, on the device the
    process group works on.  Returns the concatenation over ranks (rank order) on every rank.
    One collective for the counts, one for the padded payload.
    """
    world = dist.get_world_size(group)
    dev = u.device
    n = torch.tensor([u.numel()], dtype=torch.int64, device=dev)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, n, group=group)
    counts = counts.tolist()
    width = max(max(counts), 1)
    send = torch.zeros(3, width, dtype=torch.int32, device=dev)
    k = u.numel()
    send[0, :k] = u
    send[1, :k] = v
    send[2, :k] = s.view(torch.int32)
    recv = torch.empty(world, 3, width, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group)
    gu = torch.cat([recv[r, 0, :counts[r]] for r in range(world)])
    gv = torch.cat([recv[r, 1, :counts[r]] for r in range(world)])
    gs = torch.cat([recv[r, 2, :counts[r]] for r in range(world)]).view(torch.float32)
    return gu.contiguous(), gv.contiguous(), gs.contiguous()


def predict_distributed(pred, measure, min_degree1, max_edges, group=None, **kw):
    """One multi-GPU prediction on an already partitioned ``Predictor``.

    Returns (result-dict of the local phase, merged count, merge_ms); the merged edges are the
    handle's result afterwards (``pred.fetch``)."""
    r = pred.predict(measure, min_degree1, max_edges=max_edges, **kw)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return r, r["count"], 0.0
    n = r["count"]
    dev = torch.device("cuda", torch.cuda.current_device())
    u = torch.empty(n, dtype=torch.int32, device=dev)
    v = torch.empty(n, dtype=torch.int32, device=dev)
    s = torch.empty(n, dtype=torch.float32, device=dev)
    if n:
        pred.fetch_into(u.data_ptr(), v.data_ptr(), s.data_ptr(), n)
    gu, gv, gs = gather_candidates(u, v, s, group)
    torch.cuda.current_stream().synchronize()
    ms = pred.merge(gu.data_ptr(), gv.data_ptr(), gs.data_ptr(), gu.numel(), max_edges)
    total = min(gu.numel(), max_edges)
    return r, total, ms

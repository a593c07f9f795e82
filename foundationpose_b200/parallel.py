"""Multi-GPU register(): the hypothesis list shards contiguously over the ranks (one process per GPU).

Refinement is per-hypothesis independent (refine_network.py:73-93 has no cross-batch op in eval mode),
scoring is independent up to the per-hypothesis 512-d feature (score_network.py:60-74) and then couples
*all* hypotheses through `att_cross` (score_network.py:84-88).  So the only exchange is ONE all-gather of
[n_local, 512 + 16] floats (feature | refined pose) per rank — NCCL over NVLink/NVSwitch — after which
every rank runs the identical fp32 tail on the identical gathered buffer and obtains the same scores
and the same arg-max (bit-exact index agreement across ranks and with the single-GPU run).
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world, rank):
    """Contiguous slices whose sizes differ by at most one (252 over 8 -> 32,32,32,32,31,31,31,31)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_counts(n, world):
    return [shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world)]


_gather_ws = {}


def gather_rows(local, n_total, group=None):
    """All-gather row blocks of unequal height: local (n_local, C) -> (n_total, C) on every rank,
    rows in rank order.  One collective; shards are padded to the largest height.  The send / receive / output
    buffers and the compaction index are allocated once per (shape, world) and reused: no allocation, no torch.cat
    and no host-side loop on the per-frame path."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        assert local.shape[0] == n_total
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    C = local.shape[1]
    key = (n_total, C, world, rank, local.dtype, local.device)
    ws = _gather_ws.get(key)
    if ws is None:
        counts = shard_counts(n_total, world)
        mx = max(counts)
        idx = None
        if not all(c == mx for c in counts):
            idx = torch.cat([torch.arange(r * mx, r * mx + counts[r]) for r in range(world)]).to(local.device)
        ws = _gather_ws[key] = dict(counts=counts, mx=mx, send=local.new_zeros(mx, C), recv=local.new_empty(world * mx, C), idx=idx,
                                    out=local.new_empty(n_total, C))
    assert local.shape[0] == ws["counts"][rank], "local shard does not match shard_bounds()"
    ws["send"][: local.shape[0]].copy_(local)
    dist.all_gather_into_tensor(ws["recv"], ws["send"], group=group)
    if ws["idx"] is None:
        return ws["recv"]
    return torch.index_select(ws["recv"], 0, ws["idx"], out=ws["out"])


class ShardedRegister:
    """register() hot loop over torch.distributed: refine + featurise the local slice, all-gather once,
    replicate the cross-hypothesis tail."""

    def __init__(self, engine, group=None):
        self.engine = engine
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._packed = {}
        # the engine's tensors live on its CUDA device; a test double may say otherwise (tests/test_host_logic_cpu.py
        # runs this class over gloo with a host-side stand-in for the engine)
        self.device = torch.device(getattr(engine, "tensor_device", "cuda"))

    def run(self, poses_all, iterations):
        """poses_all: (N,4,4) identical on every rank (host or device).  Returns refined poses (N,4,4),
        scores (N,), best index — identical on every rank."""
        e = self.engine
        N = len(poses_all)
        lo, hi = shard_bounds(N, self.world, self.rank)
        local = torch.as_tensor(poses_all[lo:hi], dtype=torch.float32)
        if local.device.type != self.device.type:
            local = local.to(self.device, non_blocking=True)
        if hi > lo:
            refined, _, _ = e.refine(local, iterations)
            feats = e.score_features(refined)
            packed = self._packed.get(hi - lo)
            if packed is None:
                packed = self._packed[hi - lo] = torch.empty(hi - lo, 528, dtype=torch.float32, device=self.device)
            packed[:, :512].copy_(feats)
            packed[:, 512:].copy_(refined.reshape(-1, 16))
        else:
            packed = torch.empty(0, 528, dtype=torch.float32, device=self.device)
        allp = gather_rows(packed, N, self.group)
        scores, best = e.score_tail(allp[:, :512].contiguous())
        return allp[:, 512:].reshape(N, 4, 4), scores, best

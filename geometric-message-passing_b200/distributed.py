"""Multi-GPU for the two workload shapes of SURVEY.md §8e (the reference itself has no distributed code).

1. Batched small graphs (configs 2-4): shard by graph, no data-path collective; gradients are all-reduced after
   backward (`allreduce_gradients`), e3nn BatchNorm statistics through `gmp_b200.BatchNorm(process_group=...)`.

2. One large radius graph (config 5): partition by destination node.  Nodes are sorted along x and cut into `world`
   contiguous slabs; rank p owns the destination rows of slab p and needs, as message sources, the nodes of other slabs
   within `r` of its boundaries (the halo).  Because everything is sorted by x, what rank p sends to rank q is one
   contiguous range of its owned rows, so packing is a slice, the exchange is grouped point-to-point send/recv
   (ncclSend/ncclRecv under NCCL, also available under gloo for the CPU tests), and the backward adds the returned
   halo gradients onto those ranges in fixed rank order -- deterministic, no atomics.
   Local numbering is [left halo | owned | right halo], i.e. ascending global order, so the local radius graph lists
   every owned row's neighbours in the same order as the single-GPU graph: the partitioned reduction is the same sum.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def allreduce_gradients(params, group=None, grads=None) -> Optional[torch.Tensor]:
    """Sum the gradients of `params` across ranks through one flat buffer; writes the sums back in place.
    `grads`: the gradient tensors themselves when they do not hang off `p.grad` (GraphedStep.grads)."""
    grads = [p.grad for p in params if p.grad is not None] if grads is None else [g for g in grads if g is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    return flat


@dataclass
class SlabPartition:
    rank: int
    world: int
    n_global: int
    own_lo: int               # owned nodes are global (x-sorted) indices [own_lo, own_hi)
    own_hi: int
    n_left: int               # halo nodes with smaller / larger global index
    n_right: int
    send_ranges: List[Tuple[int, int]]   # per peer q: slice [a, b) of the OWNED rows this rank sends to q
    recv_counts: List[int]               # per peer q: rows received from q (q < rank -> left halo, q > rank -> right halo)
    local_global: torch.Tensor           # int64 [n_local]: global index of every local node (ascending)
    # peer-memory exchange (every rank computes every rank's window, so these need no communication):
    peer_send_start: Optional[List[int]] = None   # per peer q: first row, among q's OWNED rows, of the range q sends to this rank
    peer_halo_offset: Optional[List[int]] = None  # per peer q: offset, inside q's [left halo | right halo] rows, of the rows this rank owns
    n_own_max: int = 0                            # max over ranks of the owned / halo row counts (symmetric buffer sizes)
    halo_max: int = 0

    @property
    def n_own(self) -> int:
        return self.own_hi - self.own_lo

    @property
    def n_local(self) -> int:
        return self.n_left + self.n_own + self.n_right

    @property
    def own_slice(self) -> slice:
        return slice(self.n_left, self.n_left + self.n_own)


def slab_partition(x_sorted: torch.Tensor, r: float, rank: int, world: int) -> SlabPartition:
    """`x_sorted`: the x coordinates of ALL nodes in ascending order (every rank passes the same array).
    Equal-count slabs; the halo window of a slab is [x_first_owned - r, x_last_owned + r]."""
    n = x_sorted.numel()
    bounds = [(n * p) // world for p in range(world + 1)]
    # the 2 * world window edges are located on the device the array lives on (two searchsorted calls over the sorted
    # array, 2 * world values read back once); nothing of size n crosses to the host
    xs = x_sorted.detach().contiguous()   # (a column view such as pos[:, 0] is strided)
    nonempty = [p for p in range(world) if bounds[p + 1] > bounds[p]]
    wins = [(bounds[p], bounds[p]) for p in range(world)]
    if nonempty:
        first = xs[torch.tensor([bounds[p] for p in nonempty], device=xs.device)]
        last = xs[torch.tensor([bounds[p + 1] - 1 for p in nonempty], device=xs.device)]
        rw = r * (1.0 + 1e-6)        # the subtraction below rounds in the array's dtype: widen the window by more than that
        a = torch.searchsorted(xs, first - rw, right=False)
        b = torch.searchsorted(xs, last + rw, right=True)
        for p, ai, bi in zip(nonempty, a.tolist(), b.tolist()):
            wins[p] = (min(int(ai), bounds[p]), max(int(bi), bounds[p + 1]))   # a window always contains its own slab
    own_lo, own_hi = bounds[rank], bounds[rank + 1]
    a, b = wins[rank]
    n_left, n_right = own_lo - a, b - own_hi
    send, recv = [], []
    for q in range(world):
        if q == rank:
            send.append((0, 0))
            recv.append(0)
            continue
        qa, qb = wins[q]
        s_lo, s_hi = max(own_lo, qa), min(own_hi, qb)     # my owned nodes inside q's window
        send.append((s_lo - own_lo, max(s_hi, s_lo) - own_lo) if s_hi > s_lo else (0, 0))
        r_lo, r_hi = max(bounds[q], a), min(bounds[q + 1], b)  # q's owned nodes inside my window
        recv.append(max(r_hi - r_lo, 0))
    assert sum(recv[:rank]) == n_left and sum(recv[rank + 1:]) == n_right
    local_global = torch.arange(a, b, dtype=torch.int64, device=x_sorted.device)

    def recv_count(p, q):   # rows rank p receives from rank q
        return max(min(bounds[q + 1], wins[p][1]) - max(bounds[q], wins[p][0]), 0) if p != q else 0

    peer_send_start = [max(bounds[q], a) - bounds[q] if (q != rank and recv[q] > 0) else 0 for q in range(world)]
    peer_halo_offset = [sum(recv_count(q, p2) for p2 in range(rank) if p2 != q) for q in range(world)]
    n_own_max = max(bounds[p + 1] - bounds[p] for p in range(world))
    halo_max = max((bounds[p] - wins[p][0]) + (wins[p][1] - bounds[p + 1]) for p in range(world))
    return SlabPartition(rank, world, n, own_lo, own_hi, n_left, n_right, send, recv, local_global, peer_send_start,
                         peer_halo_offset, n_own_max, max(halo_max, 1))


class PeerHalo:
    """Symmetric-memory staging for the halo exchange of one partition: every rank's buffers are mapped into every other
    rank's address space (torch.distributed._symmetric_memory: CUDA VMM + NVLink P2P), so a rank PULLS its halo rows out of
    the owners' memory with its own kernel (csrc/halo.cu) after a device-side barrier -- no ncclSend / ncclRecv, no
    concatenation.  Buffers are double-buffered per (direction, row width): one barrier per exchange is enough, because a
    rank that writes buffer k again has passed the barrier of exchange k + 1, which every peer enters only after its pull
    from buffer k."""

    def __init__(self, part: SlabPartition, device, group=None):
        self.part, self.device = part, device
        self.group = dist.group.WORLD if group is None else group
        self._bufs = {}
        self._turn = {}

    def buffer(self, rows: int, feat: int, tag: str, dtype=torch.float32, copies: int = 2):
        import torch.distributed._symmetric_memory as symm
        key = (rows, feat, tag)
        if key not in self._bufs:
            pair = []
            for _ in range(copies):
                t = symm.empty(max(rows, 1) * feat, dtype=dtype, device=self.device)
                pair.append((t, symm.rendezvous(t, self.group)))
            self._bufs[key] = pair
            self._turn[key] = 0
        k = self._turn[key]
        self._turn[key] = (k + 1) % len(self._bufs[key])
        return self._bufs[key][k]


def _pull(src_ptrs, dst_ptrs, nbytes, add: bool):
    import ctypes as C
    from ._lib import call
    n = len(src_ptrs)
    call("gmp_halo_pull", (C.c_void_p * n)(*src_ptrs), (C.c_void_p * n)(*dst_ptrs), (C.c_int64 * n)(*nbytes), n, int(add))


class _PeerHaloExchange(torch.autograd.Function):
    """_HaloExchange over peer memory: forward pulls the halo rows out of the owners' staging buffers, backward lets the
    owners pull the halo gradients and add them onto the rows they had exposed, peer by peer in rank order."""

    @staticmethod
    def forward(ctx, x_own: torch.Tensor, part: SlabPartition, peer: PeerHalo):
        ctx.part, ctx.peer = part, peer
        x_own = x_own.contiguous()
        assert x_own.dtype == torch.float32 and x_own.dim() == 2
        feat = x_own.shape[1]
        rb = feat * 4
        buf, hdl = peer.buffer(part.n_own_max, feat, "fwd")
        stage = buf.view(-1, feat)
        for q in range(part.world):          # expose the boundary rows the neighbours read
            a, b = part.send_ranges[q]
            if b > a:
                stage[a:b].copy_(x_own[a:b])
        hdl.barrier()
        out = x_own.new_empty(part.n_local, feat)
        src, dst, nb, off = [], [], [], 0
        for q in range(part.world):
            if q == part.rank:
                src.append(x_own.data_ptr()); dst.append(out.data_ptr() + off * rb); nb.append(part.n_own * rb)
                off += part.n_own
                continue
            c = part.recv_counts[q]
            if c > 0:
                src.append(int(hdl.buffer_ptrs[q]) + part.peer_send_start[q] * rb); dst.append(out.data_ptr() + off * rb); nb.append(c * rb)
            off += c
        for i in range(0, len(src), 16):
            _pull(src[i:i + 16], dst[i:i + 16], nb[i:i + 16], add=False)
        return out

    @staticmethod
    def backward(ctx, g_local: torch.Tensor):
        return _return_halo_gradients(g_local, ctx.part, ctx.peer), None, None


def _return_halo_gradients(g_local: torch.Tensor, part: SlabPartition, peer: PeerHalo) -> torch.Tensor:
    """g_local [n_local, F] -> g_own [n_own, F]: the halo rows go back to their owners, which pull them out of this rank's
    staging buffer and add them onto the rows they had exposed, peer by peer in rank order."""
    g_local = g_local.contiguous()
    feat = g_local.shape[1]
    rb = feat * 4
    buf, hdl = peer.buffer(part.halo_max, feat, "bwd")
    stage = buf.view(-1, feat)
    if part.n_left:
        stage[:part.n_left].copy_(g_local[:part.n_left])
    if part.n_right:
        stage[part.n_left:part.n_left + part.n_right].copy_(g_local[part.n_left + part.n_own:])
    hdl.barrier()
    g_own = g_local[part.own_slice].clone()
    for q in range(part.world):          # one launch per peer, in rank order: overlapping ranges are added in a fixed order
        a, b = part.send_ranges[q]
        if b > a:
            _pull([int(hdl.buffer_ptrs[q]) + part.peer_halo_offset[q] * rb], [g_own.data_ptr() + a * rb], [(b - a) * rb], add=True)
    return g_own


class _PeerQFn(torch.autograd.Function):
    """Autograd edge of the fused gather: forward publishes nothing itself (the projection already wrote the owned bf16 rows
    into the peer-visible buffer) -- it orders that write before the neighbours' reads with the device-side barrier and
    returns a shape-only placeholder for `Q` over the local rows; backward receives dL/dQ for every local row (segment sums
    over the source-sorted CSR, halo rows included) and returns the halo rows' part to their owners."""

    @staticmethod
    def forward(ctx, q_own: torch.Tensor, part: SlabPartition, peer: PeerHalo, hdl):
        ctx.part, ctx.peer = part, peer
        hdl.barrier()
        return q_own.new_zeros(1, 1).expand(part.n_local, q_own.shape[1])

    @staticmethod
    def backward(ctx, g_local: torch.Tensor):
        return _return_halo_gradients(g_local, ctx.part, ctx.peer), None, None, None


class PeerQ:
    """Per layer: the peer-visible bf16 buffer of the projected rows Q = h W0b^T of the OWNED nodes (kept from the forward to
    the backward pass of the same step: the backward kernel gathers the same rows again), and the gmp_peer_rows the edge
    kernels need to find a halo source's row in the left / right neighbour's buffer.  Only neighbouring ranks may
    contribute halo rows (slabs at least one radius thick)."""

    def __init__(self, peer: PeerHalo, layer: int):
        from ._lib import PeerRows
        part = peer.part
        for q in range(part.world):
            if part.recv_counts[q] > 0 and abs(q - part.rank) != 1:
                raise NotImplementedError("fused halo gather: halo rows from non-adjacent ranks (slabs thinner than the radius)")
        self.peer, self.part = peer, part
        self.buf, self.hdl = peer.buffer(part.n_own_max, 128, f"q16.{layer}", dtype=torch.bfloat16, copies=1)
        r = part.rank
        left = (int(self.hdl.buffer_ptrs[r - 1]) + part.peer_send_start[r - 1] * 256) if part.n_left else None
        right = (int(self.hdl.buffer_ptrs[r + 1]) + part.peer_send_start[r + 1] * 256) if part.n_right else None
        self.rows = PeerRows(left, right, part.n_left, part.n_own)

    def project(self, h_own: torch.Tensor, w: torch.Tensor):
        """(Q placeholder over the local rows, bf16 owned rows in the peer-visible buffer, gmp_peer_rows)."""
        from . import nodechain as nc
        q16 = self.buf.view(-1, 128)[:self.part.n_own]
        q_own, _ = nc.ChainLinearFn.apply(h_own, w, None, True, q16)
        return _PeerQFn.apply(q_own, self.part, self.peer, self.hdl), q16, self.rows


class _HaloExchange(torch.autograd.Function):
    """owned rows [n_own, F] -> local rows [n_left + n_own + n_right, F] (halo rows filled from their owners)."""

    @staticmethod
    def forward(ctx, x_own: torch.Tensor, part: SlabPartition, group):
        ctx.part, ctx.group = part, group
        x_own = x_own.contiguous()
        F = x_own.shape[1:]
        recv = [x_own.new_empty((c,) + tuple(F)) for c in part.recv_counts]
        ops = []
        for q in range(part.world):
            a, b = part.send_ranges[q]
            if b > a:
                ops.append(dist.P2POp(dist.isend, x_own[a:b].contiguous(), q, group=group))
            if part.recv_counts[q] > 0:
                ops.append(dist.P2POp(dist.irecv, recv[q], q, group=group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return torch.cat(recv[:part.rank] + [x_own] + recv[part.rank + 1:], dim=0)

    @staticmethod
    def backward(ctx, g_local: torch.Tensor):
        part: SlabPartition = ctx.part
        g_local = g_local.contiguous()
        g_own = g_local[part.own_slice].clone()
        # halo gradients travel back to their owners, which add them onto the rows they had sent (fixed rank order)
        back = [g_own.new_empty((part.send_ranges[q][1] - part.send_ranges[q][0],) + tuple(g_own.shape[1:])) for q in range(part.world)]
        ops, off = [], 0
        for q in range(part.world):
            c = part.recv_counts[q]
            if q == part.rank:
                off += part.n_own
                continue
            if c > 0:
                ops.append(dist.P2POp(dist.isend, g_local[off:off + c].contiguous(), q, group=ctx.group))
            off += c
            if back[q].shape[0] > 0:
                ops.append(dist.P2POp(dist.irecv, back[q], q, group=ctx.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for q in range(part.world):
            a, b = part.send_ranges[q]
            if b > a:
                g_own[a:b] += back[q]
        return g_own, None, None


def halo_exchange(x_own: torch.Tensor, part: SlabPartition, group=None, peer: Optional[PeerHalo] = None) -> torch.Tensor:
    """`peer`: a PeerHalo of this partition selects the peer-memory pull kernels (CUDA, NVLink P2P); None = grouped send/recv."""
    if part.world == 1:
        return x_own
    if peer is not None:
        return _PeerHaloExchange.apply(x_own, part, peer)
    return _HaloExchange.apply(x_own, part, group)


def local_radius_graph(pos_local: torch.Tensor, r: float, part: SlabPartition, max_num_neighbors: int = 128) -> torch.Tensor:
    """Radius graph over the local node set, keeping the edges whose destination is owned (local numbering).
    Identical, row by row, to the owned rows of the single-GPU graph as long as no row is truncated."""
    from .graph import radius_graph
    ei = radius_graph(pos_local, r, None, max_num_neighbors=max_num_neighbors)
    keep = (ei[1] >= part.n_left) & (ei[1] < part.n_left + part.n_own)
    return ei[:, keep].contiguous()


class PartitionedEGNN(torch.nn.Module):
    """The layer loop of EGNNModel (models/egnn.py:71-79) on a destination-partitioned graph: per layer one halo
    exchange of (h, pos) forward and one of their gradients backward; the layer itself is the unchanged fused
    EGNNLayer applied to the local node set."""

    def __init__(self, num_layers: int = 4, emb_dim: int = 128, activation: str = "relu", aggr: str = "sum",
                 residual: bool = True, precision: str = "fp32", halo: str = "nccl"):
        """halo: "nccl" = grouped send/recv (also what the gloo CPU tests run); "peer" = pull kernels over symmetric
        memory (P2P loads over NVLink, csrc/halo.cu); "fused" (bf16 mode) = as "peer" for the positions and the returning
        gradients, while the feature rows are not exchanged at all: each rank projects Q for its owned rows into a
        peer-visible buffer and the edge kernels' gather reads halo sources' rows out of the neighbours' memory."""
        super().__init__()
        from .egnn import EGNNLayer
        assert halo in ("nccl", "peer", "fused")
        self.residual, self.halo = residual, halo
        self._peer, self._peer_q = None, None
        self.convs = torch.nn.ModuleList([EGNNLayer(emb_dim, activation, "layer", aggr, precision) for _ in range(num_layers)])

    def _peer_for(self, part: SlabPartition, device, group):
        if self.halo == "nccl" or part.world == 1:
            return None
        if self._peer is None or self._peer.part is not part:
            self._peer = PeerHalo(part, device, group)
            self._peer_q = [PeerQ(self._peer, l) for l in range(len(self.convs))] if self.halo == "fused" else None
        return self._peer

    def forward(self, h_own, pos_own, edge_index_local, part: SlabPartition, group=None):
        own = part.own_slice
        peer = self._peer_for(part, h_own.device, group)
        for l, conv in enumerate(self.convs):
            pos_loc = halo_exchange(pos_own, part, group, peer)
            if peer is not None and self._peer_q is not None:
                h_upd, pos_upd = conv(h_own, pos_loc, edge_index_local, rows=own, peer_q=self._peer_q[l])
                h_own = h_own + h_upd if self.residual else h_upd
                pos_own = pos_upd
                continue
            h_loc = halo_exchange(h_own, part, group, peer)
            h_upd, pos_upd = conv(h_loc, pos_loc, edge_index_local, rows=own)   # node-side work on owned rows only
            h_own = h_own + h_upd if self.residual else h_upd
            pos_own = pos_upd
        return h_own, pos_own

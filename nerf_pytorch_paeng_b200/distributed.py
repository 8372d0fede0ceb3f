"""Multi-GPU plumbing: one process per GPU, rays sharded across ranks.

Render needs only a final gather of the per-rank pixel bands; training is data parallel with one
all-reduce (sum) of each network's flat fp32 gradient buffer (595,844 floats per net) -- NCCL over
NVLink/NVSwitch on GPUs, gloo in the CPU tests of this host-side logic.  The reference has no
distributed code at all (main.py:166-171 is a FIXME), so there is no reference collective to mirror.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world_size):
    """Contiguous, balanced [lo, hi) band of n items for `rank` (first n % world ranks get one extra)."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class DistContext:
    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world_size = dist.get_world_size(group)

    def shard_range(self, n):
        return shard_range(n, self.rank, self.world_size)

    def allreduce_(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def allreduce_grads(self, model):
        """Sum the two flat gradient buffers over ranks.  The loss of every rank is normalised by the
        GLOBAL ray count, so the sum equals the single-GPU gradient of the concatenated batch."""
        works = []
        for net in (model.model_coarse, model.model_fine):
            g = net.bind_flat_grad()
            works.append(dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for w in works:
            w.wait()

    def allreduce_grad_async(self, net):
        """Start the sum all-reduce of one network's flat gradient; returns the work handle (wait() before the optimiser)."""
        return dist.all_reduce(net.bind_flat_grad(), op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def joint_grad_buffer(self, model):
        """One contiguous fp32 buffer [grad_coarse | grad_fine | loss_c, loss_f]; the two networks' flat gradient buffers
        become views of it, so a train step needs exactly ONE collective.  (The MLP kernels are persistent and fill every SM's
        shared memory, so a collective kernel launched "under" them does not overlap -- it runs between two of them.  Fewer
        collective launches beat earlier ones.)  Returns (whole, loss_view)."""
        nets = (model.model_coarse, model.model_fine)
        sizes = [n.flat_params().numel() for n in nets]
        buf = getattr(self, '_joint', None)
        if buf is None or buf.numel() != sum(sizes) + 2 or buf.device != nets[0].flat.device:
            buf = self._joint = torch.zeros(sum(sizes) + 2, dtype=torch.float32, device=nets[0].flat.device)
        off = 0
        for n, sz in zip(nets, sizes):
            view = buf[off:off + sz]
            if n.flat_grad is None or n.flat_grad.data_ptr() != view.data_ptr():
                n.flat_grad = view
            n.bind_flat_grad()
            off += sz
        return buf, buf[off:off + 2]

    def gather_rows(self, local, n_total):
        """All-gather variable-length row bands (dim 0) into the full [n_total, ...] tensor."""
        sizes = [shard_range(n_total, r, self.world_size) for r in range(self.world_size)]
        maxlen = max(hi - lo for lo, hi in sizes)
        pad = torch.zeros((maxlen,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
        bufs = [torch.empty_like(pad) for _ in range(self.world_size)]
        dist.all_gather(bufs, pad, group=self.group)
        return torch.cat([b[:hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)


class PeerGradExchange:
    """Gradient exchange without a collective kernel: every rank PUSHES its flat gradient into
    a slot of every peer's symmetric-memory buffer with plain device-to-device copies -- executed by the copy engines over
    NVLink, so they run UNDER the persistent MLP kernels that own every SM (an NCCL kernel cannot: it waits for an SM) -- and
    the sum over ranks is folded into Adam's load (nb_adam_step_sum, rank-ordered => bit-identical replicas).

    Schedule of one step: the coarse network's backward finishes before the fine pass starts (nb_train_rays nets=1 then 2), so
    its 2.4 MB gradient travels during the fine forward/backward; only the fine network's push and one cross-rank barrier are
    exposed at the end of the step.  Receive buffers are double-buffered by step parity."""

    def __init__(self, ctx, model):
        import torch.distributed._symmetric_memory as symm
        self.ctx = ctx
        self.world, self.rank = ctx.world_size, ctx.rank
        nets = (model.model_coarse, model.model_fine)
        self.sizes = [n.flat_params().numel() for n in nets]
        self.G = sum(self.sizes) + 2                                  # + the two losses
        dev = nets[0].flat.device
        self.buf = symm.empty(2 * self.world * self.G, dtype=torch.float32, device=dev)
        self.buf.zero_()
        group = ctx.group if ctx.group is not None else dist.group.WORLD
        self.hdl = symm.rendezvous(self.buf, group.group_name if hasattr(group, 'group_name') else group)
        self.peer = [self.hdl.get_buffer(r, (2, self.world, self.G), torch.float32) for r in range(self.world)]
        self.local = self.peer[self.rank]
        self.whole, self.loss = ctx.joint_grad_buffer(model)          # [grad_c | grad_f | loss_c, loss_f]
        self.sides = [torch.cuda.Stream(device=dev) for _ in range(max(1, min(4, self.world - 1)))]   # several copy engines at once
        self.side = self.sides[0]
        self.step_parity = 0
        torch.cuda.synchronize(dev)
        self.hdl.barrier(channel=0)

    def push(self, lo, hi):
        """Start copying whole[lo:hi] into slot `rank` of every peer (side stream, after everything enqueued so far)."""
        ev = torch.cuda.Event()
        ev.record()
        for st in self.sides:
            st.wait_event(ev)
        for k in range(1, self.world):
            r = (self.rank + k) % self.world                          # staggered targets: no two ranks start on the same peer
            with torch.cuda.stream(self.sides[(k - 1) % len(self.sides)]):
                self.peer[r][self.step_parity, self.rank, lo:hi].copy_(self.whole[lo:hi], non_blocking=True)

    def finish(self):
        """All pushes of this step have landed everywhere: barrier on the side stream, then the main stream waits for it.
        Returns the rank-ordered list of gradient sources for nb_adam_step_sum (own buffer at own rank)."""
        for st in self.sides[1:]:
            ev = torch.cuda.Event()
            ev.record(st)
            self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            self.hdl.barrier(channel=0)
            done = torch.cuda.Event()
            done.record()
        torch.cuda.current_stream().wait_event(done)
        par = self.step_parity
        self.step_parity ^= 1
        return [self.whole if r == self.rank else self.local[par, r] for r in range(self.world)]


def init_from_env(backend=None):
    """torchrun-style init (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT)."""
    import os
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world <= 1:
        return None
    if not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if torch.cuda.is_available():
            torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group(backend=backend)
    return DistContext()

"""CPU: host-side logic -- config parsing, state_dict compatibility, ray sharding, and the N>1 path
(world_size 2 over gloo)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden


def test_config_defaults_and_files(tmp_path):
    from nerf_pytorch_paeng_b200.config import get_args_parser
    cfg = tmp_path / 'lego.txt'
    cfg.write_text('# >> Setting\ngpu_ids = [1]\ndata_type = blender\nnear = 2.\nfar = 6.\nbkg_white_true\n'
                   'global_batch_false\nN_rays = 4096   # rays\nidx_save = 100000\niter_N = 200000\nexp_name = blender_lego\n')
    o = get_args_parser(['--config', str(cfg)])
    assert (o.near, o.far, o.gpu_ids, o.world_size, o.rank) == (2., 6., [1], 1, 0)
    assert o.bkg_white is True and o.global_batch is False
    # defaults of config.py:54-76
    assert (o.L_x, o.L_d, o.netDepth, o.netWidth) == (10, 4, 8, 256)
    assert (o.N_rays, o.N_samples_c, o.N_samples_f, o.chunk_rays, o.chunk_pts) == (4096, 64, 128, 4096, 524288)
    assert (o.lr, o.lr_min, o.iter_warmup, o.precrop_iters) == (5e-4, 5e-5, 10000, 0)
    o2 = get_args_parser(['--config', str(cfg), '--N_rays', '1024', '--gpu_ids', '0', '1'])
    assert o2.N_rays == 1024 and o2.gpu_ids == [0, 1] and o2.world_size == 2


def test_state_dict_matches_reference_layout():
    from nerf_pytorch_paeng_b200.model import NeRF
    g = load_golden('mlp_w256_seed0.npz')
    torch.manual_seed(0)
    net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(np.eye(3), np.zeros((2, 4, 4))))
    sd = net.state_dict()
    assert list(sd.keys()) == [str(s) for s in g['param_names']]
    np.testing.assert_allclose([float(v.double().sum()) for v in sd.values()], g['param_sums'], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose([float(v.flatten()[0]) for v in sd.values()], g['param_first'], rtol=0, atol=0)
    K, E = net.get_camera_gt()
    assert K.shape == (3, 3) and E.shape == (2, 4, 4)
    assert sd['model_coarse.linear_x.5.weight'].shape == (256, 319)
    assert sd['model_fine.linear_d.weight'].shape == (128, 283)


def test_shard_range_partitions():
    from nerf_pytorch_paeng_b200.distributed import shard_range
    for n in (0, 1, 7, 640000, 762048):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_ray_batch_cursor():
    from nerf_pytorch_paeng_b200.utils import GetterRayBatchIdx
    table = torch.arange(10 * 9, dtype=torch.float32).reshape(10, 3, 3)
    g = GetterRayBatchIdx(table)
    i, t, e = g(4)
    assert (i, e) == (4, 0) and torch.equal(t, table)
    i, t, e = g(4)
    assert (i, e) == (8, 0)
    i, t, e = g(4)               # 12 >= 10 -> reshuffle (utils.py:54-58)
    assert (i, e) == (4, 1) and sorted(t[:, 0, 0].tolist()) == sorted(table[:, 0, 0].tolist())


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    from nerf_pytorch_paeng_b200 import distributed
    dist.init_process_group('gloo', rank=rank, world_size=world)
    ctx = distributed.DistContext()

    class Net:
        def __init__(self, v):
            self.g = torch.full((11,), float(v))

        def bind_flat_grad(self):
            return self.g

    class Model:
        pass
    m = Model()
    m.model_coarse, m.model_fine = Net(rank + 1), Net(10 * (rank + 1))
    ctx.allreduce_grads(m)                               # sum over ranks
    n = 13
    lo, hi = ctx.shard_range(n)
    local = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 3)
    full = ctx.gather_rows(local, n)
    t = ctx.allreduce_(torch.tensor([float(rank)]))
    torch.save({'gc': m.model_coarse.g, 'gf': m.model_fine.g, 'full': full, 't': t}, os.path.join(tmp, f'r{rank}.pt'))
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    """The multi-GPU host logic on CPU: gradient all-reduce (sum), ragged row all-gather, rank bands."""
    world, port = 2, 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        d = torch.load(os.path.join(tmp_path, f'r{r}.pt'))
        assert torch.equal(d['gc'], torch.full((11,), 3.0)) and torch.equal(d['gf'], torch.full((11,), 30.0))
        assert torch.equal(d['full'], torch.arange(13, dtype=torch.float32)[:, None].repeat(1, 3))
        assert float(d['t']) == 1.0


def test_flat_adam_is_a_torch_optimizer_with_adam_state_dict():
    """FlatAdam vs the reference's optimizer contract (main.py:79-90,111-115; scheduler.py:6): a checkpoint WRITTEN BY THE
    REFERENCE (tests/golden/ref_checkpoint_w64_3.pth.tar, oracle/make_golden.py::gen_checkpoint) loads, round-trips in
    torch.optim.Adam's format, and LR schedulers accept the optimizer.  Host-side logic only (no compute)."""
    import os
    from conftest import GOLDEN
    from nerf_pytorch_paeng_b200 import trainer
    from nerf_pytorch_paeng_b200.model import NeRF
    ck = torch.load(os.path.join(GOLDEN, 'ref_checkpoint_w64_3.pth.tar'), map_location='cpu')
    assert set(ck) == {'idx', 'model_state_dict', 'optimizer_state_dict'} and ck['idx'] == 3       # train.py:107-109
    net = NeRF(8, 64, 63, 27, [4], gt_camera_param=(None, None))
    net.load_state_dict(ck['model_state_dict'])                                                    # strict: same keys / shapes
    opt = trainer.FlatAdam(net, lr=5e-4)
    assert isinstance(opt, torch.optim.Optimizer) and len(opt.param_groups) == 1 and len(opt.param_groups[0]['params']) == 48
    opt.load_state_dict(ck['optimizer_state_dict'])
    ref = ck['optimizer_state_dict']
    assert opt.step_count == 3 and opt.param_groups[0]['lr'] == ref['param_groups'][0]['lr']
    sd = opt.state_dict()
    assert sd['param_groups'][0]['params'] == ref['param_groups'][0]['params'] and set(sd['state']) == set(ref['state'])
    for i, st in ref['state'].items():
        assert float(sd['state'][i]['step']) == float(st['step'])
        assert torch.equal(sd['state'][i]['exp_avg'], st['exp_avg']) and torch.equal(sd['state'][i]['exp_avg_sq'], st['exp_avg_sq'])
    # the moments are views of one flat buffer per network, in parameters() order
    m, v = opt._flat_state[id(net.model_coarse)]
    p0 = net.model_coarse._plist[0]
    assert opt.state[p0]['exp_avg'].data_ptr() == m.data_ptr()
    # and the other way: torch's own Adam accepts a FlatAdam state_dict (a reference main.py can resume our checkpoints)
    a = torch.optim.Adam(net.parameters(), lr=5e-4, betas=(0.9, 0.999))
    a.load_state_dict(sd)
    assert torch.equal(a.state[p0]['exp_avg'], ref['state'][0]['exp_avg'])
    # schedulers: torch's, and the reference's own CosineAnnealingWarmupRestarts when baseline/_ref is installed
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda e: 0.5)
    assert abs(opt.param_groups[0]['lr'] - 0.5 * sched.base_lrs[0]) < 1e-12
    from baseline import ref_shim
    if ref_shim.available():
        cos = ref_shim.import_reference().scheduler.CosineAnnealingWarmupRestarts(opt, first_cycle_steps=200001, cycle_mult=1., max_lr=5e-4,
                                                                                  min_lr=5e-5, warmup_steps=10000)
        assert abs(opt.param_groups[0]['lr'] - 5e-5) < 1e-12                                       # scheduler.py:48-52 init_lr
        assert cos.optimizer is opt


def test_wgrad_claim_protocol_model():
    """Model of the work distribution of mlp_wgrad_kernel (csrc/nb_mlp_tc_bwd.cu, producer thread): per-job unit counters, home
    job by byte share, claims of up to `chunk` units that shrink towards the end of a job, early claim of the next chunk, move to
    the job with the most unclaimed bytes.  Under random interleavings of the CTAs' atomic operations every unit of every job is
    processed exactly once, every CTA terminates, and a CTA flushes its accumulator (ends a segment) exactly when it leaves a job."""
    import random

    def run(n_units, weights, grid, chunk, seed):
        rnd = random.Random(seed)
        n_jobs = len(weights)
        counters = [0] * n_jobs
        work_begin, total = [], 0
        for w in weights:
            work_begin.append(total)
            total += w * n_units
        share = [max(1, grid * w * n_units // total) for w in weights]
        sizes = []

        def claim(j, seen):
            k = max(1, min(chunk, -(-(n_units - seen) // (2 * share[j]))))
            ua = counters[j]
            counters[j] += k
            sizes.append(k)
            return ua < n_units, ua, min(ua + k, n_units)

        def cta(b):
            """generator: yields at every global-memory operation (the scheduler interleaves there); returns its segments"""
            lo = total * b // grid
            j = max(k for k in range(n_jobs) if work_begin[k] <= lo)
            segments = []
            while True:
                yield
                seen = counters[j]
                yield
                got, ua, ub = claim(j, seen)
                while not got:
                    yield
                    left = [(n_units - counters[k]) * weights[k] for k in range(n_jobs)]      # the scan (plain loads)
                    best = max(left)
                    if best <= 0:
                        return segments
                    j = left.index(best)
                    seen = counters[j]
                    yield
                    got, ua, ub = claim(j, seen)
                seg = []
                while got:
                    yield
                    more, na, nb = claim(j, ub)           # next chunk of the same job, claimed before this one is streamed
                    seg.extend((j, u) for u in range(ua, ub))
                    got, ua, ub = more, na, nb
                segments.append(seg)                      # `last` flag on the final unit: accumulator flushed here

        gens = {b: cta(b) for b in range(grid)}
        done = {}
        steps = 0
        while gens:
            b = rnd.choice(list(gens))
            try:
                next(gens[b])
            except StopIteration as e:
                done[b] = e.value
                del gens[b]
            steps += 1
            assert steps < 10_000_000, 'claim protocol does not terminate'
        seen_units = {}
        for b, segs in done.items():
            for seg in segs:
                assert seg and len({j for j, _ in seg}) == 1          # a segment is one job
                for ju in seg:
                    assert ju not in seen_units, ('claimed twice', ju)
                    seen_units[ju] = b
        assert len(seen_units) == n_jobs * n_units
        assert max(counters) <= n_units + chunk * (2 * grid + n_jobs)          # over-claims stay bounded (uint32 counters)
        assert max(sizes) <= chunk and min(sizes) >= 1
        return done

    weights = [5, 8, 8, 8, 8, 5, 8, 8, 8, 7, 4]                        # operand blobs per unit of the 11 jobs
    for n_units, grid, chunk, seed in [(6, 6, 1, 0), (4, 4, 1, 1), (97, 13, 4, 2), (900, 148, 4, 3), (4096, 148, 8, 4), (12288, 148, 16, 5),
                                       (2, 2, 1, 6), (33, 148, 1, 7)]:
        done = run(n_units, weights, min(grid, n_units), chunk, seed)
        assert len(done) == min(grid, n_units)


def test_wgrad_ring_parity_model():
    """Model of mlp_wgrad_kernel's operand ring: up to six stage barriers, each with its OWN phase parity kept in a bit mask by the
    producer (initially all ones: the first wait on a fresh barrier passes) and by every consumer (initially zero), because the number
    of stages in use changes from segment to segment; the ring is drained between segments and stage numbering restarts at 0.  Under
    random scheduling: the consumer reads exactly the sequence the producer wrote, no stage is refilled before it was released, and
    nobody deadlocks."""
    import random

    class Bar:                                   # mbarrier with a fixed arrival count; try_wait.parity semantics
        def __init__(self, count):
            self.count, self.pending, self.phase = count, count, 0

        def arrive(self):
            self.pending -= 1
            if self.pending == 0:
                self.pending, self.phase = self.count, self.phase + 1

        def done(self, parity):                  # "the phase with this parity has completed"
            return (self.phase & 1) != parity

    def run(segments, seed):
        rnd = random.Random(seed)
        full = [Bar(1) for _ in range(6)]
        empty = [Bar(2) for _ in range(6)]       # two consumers (MMA commit + the column-sum warps, as one party each)
        slots = [None] * 6
        written, read = [], [[], []]
        END = ('end',)

        def producer():
            pmask, any_seg = 0x3F, False
            for ns, items in segments:
                if any_seg:
                    for i in range(6):           # drain
                        while not empty[i].done((pmask >> i) & 1):
                            yield
                any_seg = True
                stage = 0
                for k, it in enumerate(items):
                    while not empty[stage].done((pmask >> stage) & 1):
                        yield
                    pmask ^= 1 << stage
                    assert slots[stage] is None, 'stage refilled before it was released'
                    slots[stage] = (it, ns, k == len(items) - 1)
                    written.append(it)
                    full[stage].arrive()
                    yield
                    stage = (stage + 1) % ns
            if any_seg:
                for i in range(6):
                    while not empty[i].done((pmask >> i) & 1):
                        yield
            slots[0] = (END, 1, True)
            full[0].arrive()

        def consumer(who):
            cmask, stage = 0, 0
            while True:
                while not full[stage].done((cmask >> stage) & 1):
                    yield
                cmask ^= 1 << stage
                it, ns, last = slots[stage]
                if it is END:
                    return
                read[who].append(it)
                yield
                empty[stage].arrive()
                if empty[stage].pending == empty[stage].count:      # both consumers released it
                    slots[stage] = None
                stage = 0 if last else (stage + 1) % ns

        gens = [producer(), consumer(0), consumer(1)]
        alive = list(range(3))
        steps = 0
        while alive:
            i = rnd.choice(alive)
            try:
                next(gens[i])
            except StopIteration:
                alive.remove(i)
            steps += 1
            assert steps < 2_000_000, 'ring protocol deadlocks'
        assert read[0] == written and read[1] == written

    rs = random.Random(0)
    for seed in range(20):
        segs, n = [], 0
        for _ in range(rs.randint(0, 4)):
            ns, cnt = rs.choice([3, 3, 5, 6]), rs.randint(1, 40)
            segs.append((ns, list(range(n, n + cnt))))
            n += cnt
        run(segs, seed)

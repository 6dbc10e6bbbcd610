"""The caller of the hot path: the geometric attack loop, restated in PyTorch on top of
``nn_distance`` (BASELINE config 3).

Reference (TensorFlow 1.13, cannot run here): ``src/adv_ae.py`` -- graph :78-153, loop :191-251,
``src/adversary.py`` (perturbation variable), ``src/encoders_decoders.py`` +
``src/ae_templates.py:22-31`` (PointNet auto-encoder).  The AE layers are ordinary PyTorch
ops (library code, outside the hot path); the two Chamfer terms per step are the library's
CUDA kernels.  One attack iteration = everything ``adv_ae.py:217-246`` does once:

    update      adv = x + pert;  recon = AE(adv)
                loss = sum_b( CD(recon, target) + w * CD(adv, x) )            (:105,120-121,131-133)
                Adam step on pert only                                          (:152-153)
    metrics     the six loss vectors re-evaluated AFTER the update              (:219-221)
    best-so-far from iteration 400 on, per example, keep the adversarial cloud with the lowest
                target reconstruction error                                     (:234-246)

The whole iteration is captured in a CUDA graph (our kernels launch on the capture stream and
allocate nothing), so 500 iterations are 500 graph replays with no host round trip; the
reference does >= 2 ``sess.run`` calls with numpy feeds per iteration.

Deliberate, documented differences:
* Adam's moment slots are reset whenever the perturbation is re-initialised.  The reference
  creates them once per class and never resets them (adv_ae.py:152 vs :214), which makes
  a batch depend on all batches before it; resetting makes pairs independent, so that pairs
  can be sharded over GPUs with bit-identical results.  ``reset_adam=False`` restores the quirk.
* single_forward=True (default): the reference evaluates the six loss vectors in a second sess.run AFTER the update
  (adv_ae.py:219-221), i.e. a second AE forward and two more Chamfer searches per iteration.  Those values are exactly
  what the forward pass of the NEXT iteration computes (same perturbation, same frozen network), so the iteration
  here is forward -> best-so-far bookkeeping for the previous update -> backward -> Adam, plus one closing forward
  after the last update: same metrics for every update, one AE forward and two searches per iteration instead of
  two and four.  single_forward=False keeps the reference's order of evaluation (tests compare the two).
* Update rule follows TensorFlow's Adam (epsilon outside the bias correction:
  lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps)), not torch.optim.Adam.
"""
import math

import torch
from torch import nn

from . import ops


class PointNetAE(nn.Module):
    """Victim auto-encoder of ae_templates.py:22-31: encoder conv1d(k=1) 3-64-128-128-256-bneck with
    BatchNorm + ReLU and a max over points (encoders_decoders.py:37-72), decoder FC
    bneck-256-256-(N*3) with ReLU between layers, none at the end (:100-132)."""

    def __init__(self, n_points=2048, bneck=128):
        super().__init__()
        self.n_points = n_points
        chans = [3, 64, 128, 128, 256, bneck]
        layers = []
        for cin, cout in zip(chans[:-1], chans[1:]):
            layers += [nn.Conv1d(cin, cout, 1), nn.BatchNorm1d(cout, momentum=0.1), nn.ReLU(inplace=True)]
        self.encoder = nn.Sequential(*layers)
        self.decoder = nn.Sequential(nn.Linear(bneck, 256), nn.ReLU(inplace=True), nn.Linear(256, 256),
                                     nn.ReLU(inplace=True), nn.Linear(256, n_points * 3))

    def encode(self, pc):
        return self.encoder(pc.transpose(1, 2)).amax(dim=2)

    def forward(self, pc):
        z = self.encode(pc)
        return self.decoder(z).view(-1, self.n_points, 3), z


def chamfer_per_pc(a, b, fused=True):
    """mean(d_ab,1) + mean(d_ba,1) and max(d_ab,1), as adv_ae.py:120-121,131-133 build them: one search and one
    reduction launch (ops.chamfer_loss_terms); fused=False builds them from nn_distance with torch reductions."""
    if fused and a.is_cuda:
        cd, mx, _, _ = ops.chamfer_loss_terms(a, b)
        return cd, mx
    d1, _, d2, _ = ops.nn_distance(a, b)
    return d1.mean(dim=1) + d2.mean(dim=1), d1.amax(dim=1)


class GeometricAttack:
    """Output-space geometric attack (runner_attacker.sh:7: loss_adv_type=chamfer,
    loss_dist_type=chamfer, dist_weight 1.0) for one batch of (source, target) pairs."""

    def __init__(self, ae, batch_size, n_points=2048, lr=0.01, dist_weight=1.0, num_iterations=500,
                 num_iterations_thresh=400, use_cuda_graph=True, reset_adam=True, device="cuda", single_forward=True,
                 fused_loss=True):
        self.ae = ae.to(device).eval()  # is_training(False): BatchNorm frozen (adv_ae.py:210)
        for p in self.ae.parameters():
            p.requires_grad_(False)
        self.B, self.N = batch_size, n_points
        self.lr, self.w = lr, float(dist_weight)
        self.iters, self.thresh = num_iterations, num_iterations_thresh
        self.device = torch.device(device)
        self.reset_adam = reset_adam
        self.single_forward = single_forward
        self.fused_loss = fused_loss
        self.use_graph = use_cuda_graph and self.device.type == "cuda"
        f32 = dict(dtype=torch.float32, device=self.device)
        self.x = torch.zeros(batch_size, n_points, 3, **f32)        # source clouds
        self.gt = torch.zeros(batch_size, n_points, 3, **f32)       # target clouds
        self.ref = torch.ones(batch_size, **f32)                    # target_ae_loss_ref
        self.pert = torch.zeros(batch_size, n_points, 3, **f32).requires_grad_(True)
        self.m = torch.zeros_like(self.pert)
        self.v = torch.zeros_like(self.pert)
        self.t = torch.zeros((), **f32)
        self.collect = torch.zeros((), dtype=torch.bool, device=self.device)  # iteration+1 >= thresh
        self.collect_prev = torch.zeros((), dtype=torch.bool, device=self.device)  # the same for the previous update
        self.metrics_graph = None
        self.best_err = torch.full((batch_size,), 1e10, **f32)
        self.best_metrics = torch.zeros(batch_size, 4, **f32)       # loss_adv, loss_dist, source CD, target NRE
        self.best_adv = torch.zeros(batch_size, n_points, 3, **f32)
        self.best_recon = torch.zeros(batch_size, n_points, 3, **f32)
        self.last = {}
        self.graph = None

    # -- one iteration (adv_ae.py:217-246) ------------------------------------------------
    def _bookkeeping(self, adv, recon, err, src_cd, src_max, collect):
        """Metrics of the current perturbation and the best-so-far update (adv_ae.py:219-246)."""
        pert_sq = (self.pert * self.pert).sum(dim=2)
        self.last = {"loss_adv": err, "loss_dist": src_cd, "loss_pert": pert_sq.sum(dim=1).sqrt(),
                     "loss_max": src_max, "source_chamfer_dist": src_cd, "target_recon_error": err}
        better = collect & (err < self.best_err)
        self.best_err.copy_(torch.where(better, err, self.best_err))
        met = torch.stack([err, src_cd, src_cd, err / self.ref], dim=1)
        self.best_metrics.copy_(torch.where(better[:, None], met, self.best_metrics))
        self.best_adv.copy_(torch.where(better[:, None, None], adv, self.best_adv))
        self.best_recon.copy_(torch.where(better[:, None, None], recon, self.best_recon))

    def _adam(self, g):
        b1, b2, eps = 0.9, 0.999, 1e-8
        self.t += 1.0
        self.m.mul_(b1).add_(g, alpha=1 - b1)
        self.v.mul_(b2).addcmul_(g, g, value=1 - b2)
        lr_t = self.lr * torch.sqrt(1 - b2 ** self.t) / (1 - b1 ** self.t)
        self.pert.sub_(lr_t * self.m / (self.v.sqrt() + eps))

    def _iteration(self):
        if self.single_forward:
            return self._iteration_single()
        adv = self.x + self.pert
        recon, _ = self.ae(adv)
        loss_adv, _ = chamfer_per_pc(recon, self.gt, self.fused_loss)
        loss_dist, _ = chamfer_per_pc(adv, self.x, self.fused_loss)
        loss = (loss_adv + self.w * loss_dist).sum()
        (g,) = torch.autograd.grad(loss, self.pert)
        with torch.no_grad():
            self._adam(g)
            # second sess.run: metrics at the UPDATED perturbation
            adv = self.x + self.pert
            recon, _ = self.ae(adv)
            err, _ = chamfer_per_pc(recon, self.gt, self.fused_loss)            # loss_ae_per_pc == loss_adv
            src_cd, src_max = chamfer_per_pc(adv, self.x, self.fused_loss)      # input_dist_per_pc, max_dist_per_pc
            self._bookkeeping(adv, recon, err, src_cd, src_max, self.collect)

    def _iteration_single(self):
        """forward at the current perturbation = metrics of the PREVIOUS update (collect_prev) + loss of this one."""
        adv = self.x + self.pert
        recon, _ = self.ae(adv)
        loss_adv, _ = chamfer_per_pc(recon, self.gt, self.fused_loss)
        loss_dist, src_max = chamfer_per_pc(adv, self.x, self.fused_loss)
        with torch.no_grad():
            self._bookkeeping(adv.detach(), recon.detach(), loss_adv.detach(), loss_dist.detach(), src_max.detach(),
                              self.collect_prev)
        loss = (loss_adv + self.w * loss_dist).sum()
        (g,) = torch.autograd.grad(loss, self.pert)
        with torch.no_grad():
            self._adam(g)

    def _metrics_only(self):
        """Closing forward of the single-forward mode: the metrics of the last update."""
        with torch.no_grad():
            adv = self.x + self.pert
            recon, _ = self.ae(adv)
            err, _ = chamfer_per_pc(recon, self.gt, self.fused_loss)
            src_cd, src_max = chamfer_per_pc(adv, self.x, self.fused_loss)
            self._bookkeeping(adv, recon, err, src_cd, src_max, self.collect_prev)

    def _capture(self):
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(2):  # warm-up outside capture (cuDNN/cuBLAS workspaces, our smem attributes)
                self._iteration()
        torch.cuda.current_stream(self.device).wait_stream(s)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._iteration()

    def init_pert(self, stddev=1e-7, seed=55, pair_ids=None):
        """adversary.py:27-28: truncated normal, sigma 1e-7, fixed seed.  The reference draws one
        (B,N,3) tensor per batch, so a pair's noise depends on its position in the batch; with
        `pair_ids` every pair gets its own stream (seed, pair id) and the result no longer depends
        on how pairs are batched or sharded over GPUs."""
        p = torch.empty(self.B, self.N, 3)
        if pair_ids is None:
            g = torch.Generator(device="cpu").manual_seed(seed)
            nn.init.trunc_normal_(p, mean=0.0, std=stddev, a=-2 * stddev, b=2 * stddev, generator=g)
        else:
            for row, pid in enumerate(pair_ids):
                g = torch.Generator(device="cpu").manual_seed(seed * 1000003 + int(pid))
                nn.init.trunc_normal_(p[row], mean=0.0, std=stddev, a=-2 * stddev, b=2 * stddev, generator=g)
        with torch.no_grad():
            self.pert.copy_(p.to(self.device))
            if self.reset_adam:
                self.m.zero_()
                self.v.zero_()
                self.t.zero_()

    def step(self):
        if self.use_graph:
            if self.graph is None:
                state = [t.clone() for t in (self.pert.detach(), self.m, self.v, self.t, self.best_err,
                                             self.best_metrics, self.best_adv, self.best_recon)]
                self._capture()
                with torch.no_grad():  # undo the warm-up iterations
                    for dst, src in zip((self.pert, self.m, self.v, self.t, self.best_err, self.best_metrics,
                                         self.best_adv, self.best_recon), state):
                        dst.copy_(src)
            self.graph.replay()
        else:
            self._iteration()

    def run(self, source_pc, target_pc, target_ae_loss_ref=None, iterations=None, pair_ids=None):
        """_attack_one_batch for one dist_weight: returns metrics (B,5) [loss_adv, loss_dist,
        source_chamfer_dist, target_nre, target_recon_error], adversarial inputs, reconstructions."""
        iters = self.iters if iterations is None else iterations
        with torch.no_grad():
            self.x.copy_(source_pc)
            self.gt.copy_(target_pc)
            if target_ae_loss_ref is not None:
                self.ref.copy_(target_ae_loss_ref)
            else:
                self.ref.fill_(1.0)
            self.best_err.fill_(1e10)
            self.best_metrics.zero_()
            self.best_adv.zero_()
            self.best_recon.zero_()
        self.init_pert(pair_ids=pair_ids)
        for it in range(iters):
            self.collect.fill_((it + 1) >= min(self.thresh, iters))
            self.collect_prev.fill_(it >= 1 and it >= min(self.thresh, iters))  # update `it` is iteration it - 1
            self.step()
        if self.single_forward:
            self.collect_prev.fill_(iters >= 1 and iters >= min(self.thresh, iters))
            self._metrics_only()
        metrics = torch.cat([self.best_metrics, self.best_err[:, None]], dim=1)
        return metrics.clone(), self.best_adv.clone(), self.best_recon.clone()


def attack_pair_range(ae, sources, targets, lo, hi, batch_size=10, **kw):
    """Attack the pairs with global indices [lo, hi) in batches of `batch_size`."""
    dev = kw.get("device", "cuda")
    src, tgt = sources[lo:hi], targets[lo:hi]
    n = hi - lo
    atk = GeometricAttack(ae, batch_size, n_points=sources.shape[1], **kw)
    mets, advs = [], []
    for s in range(0, n, batch_size):
        e = min(n, s + batch_size)
        sb, tb = src[s:e].to(dev), tgt[s:e].to(dev)
        ids = list(range(lo + s, lo + e))
        if e - s < batch_size:  # pad the last batch; padded rows are dropped below
            padn = batch_size - (e - s)
            sb = torch.cat([sb, sb[-1:].expand(padn, -1, -1)])
            tb = torch.cat([tb, tb[-1:].expand(padn, -1, -1)])
            ids += [ids[-1]] * padn
        m, a, _ = atk.run(sb, tb, pair_ids=ids)
        mets.append(m[: e - s])
        advs.append(a[: e - s])
    mets = torch.cat(mets) if mets else torch.zeros(0, 5, device=dev)
    advs = torch.cat(advs) if advs else torch.zeros(0, sources.shape[1], 3, device=dev)
    return mets, advs


def attack_pairs(ae, sources, targets, batch_size=10, group=None, **kw):
    """Attack every (source, target) pair, pairs sharded contiguously over the ranks of `group`
    and the per-pair results all-gathered at the end (6 MB for 250 pairs)."""
    from . import sharding
    _, _, (lo, hi) = sharding.shard_pairs(sources, targets, group)
    mets, advs = attack_pair_range(ae, sources, targets, lo, hi, batch_size, **kw)
    total = sources.shape[0]
    return sharding.all_gather_rows(mets, total, group), sharding.all_gather_rows(advs, total, group)


def steps_per_second(batch_size=50, n_points=2048, iters=50, warmup=10, use_cuda_graph=True, device="cuda",
                     seed=0):
    """Benchmark helper: attack iterations per second for a random-init AE and synthetic clouds."""
    torch.manual_seed(seed)
    ae = PointNetAE(n_points)
    atk = GeometricAttack(ae, batch_size, n_points, use_cuda_graph=use_cuda_graph, device=device)
    g = torch.Generator().manual_seed(seed)
    atk.x.copy_((torch.rand(batch_size, n_points, 3, generator=g) - 0.5).to(device))
    atk.gt.copy_((torch.rand(batch_size, n_points, 3, generator=g) - 0.5).to(device))
    atk.init_pert()
    for _ in range(warmup):
        atk.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        atk.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return 1000.0 / ms, ms, math.nan if not atk.last else float(atk.last["loss_adv"].mean())

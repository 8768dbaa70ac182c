"""View-sharded data parallelism for one optimisation step (new functionality: the reference has no distributed path
that runs).

Views of a step are independent given the replicated scene parameters, so they shard with NO data-path collective:
rank r renders a contiguous range of the B views and back-propagates into its own full-size parameter gradients;
ONE all-reduce(SUM) over a single flat fp32 bucket (about 9.4 MB for 10 blocks with 256^2 textures) over
NVLink 5 / NVSwitch then gives every rank the full gradient.  Correctness conditions handled here:
  * the RGB loss is a MEAN over all B views (dbw.py:367): each rank divides by the GLOBAL pixel count
    (model.n_total_views), so the SUM of the per-rank gradients is the gradient of the global mean;
  * view-independent terms (parsimony / TV / overlap, dbw.py:373-405) are computed identically on every rank and
    divided by the world size before the SUM;
  * the opacity noise (dbw.py:300-301) and the overlap sample points (dbw.py:393) come from a generator that every
    rank seeds identically at every step."""
import torch
import torch.distributed as dist


def shard_views(n_views, world_size, rank):
    """Contiguous, balanced view ranges: the first n_views % world_size ranks get one extra view (49 -> 7,6,6,6,6,6,6,6)."""
    base, extra = divmod(n_views, world_size)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


class GradBucket:
    """All parameter gradients as views into one flat buffer -> a single all-reduce per step."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(n, dtype=ref.dtype, device=ref.device)
        self._zeros = {}
        off, self.views = 0, []
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            p.grad = self.views[-1]
            off += p.numel()

    def zero_(self):
        self.flat.zero_()

    def backward(self, total):
        """Back-propagate `total` and leave every parameter's gradient in the flat bucket with ONE gather kernel: the
        parameters enter the backward without .grad, so autograd hands over each gradient tensor as it is (no zero-fill
        of the bucket, no accumulate-add per parameter); a single cat then writes them into the flat buffer and the
        .grad attributes become the bucket views again.  CUDA-graph capturable."""
        for p in self.params:
            p.grad = None
        total.backward()
        pieces = []
        for p in self.params:
            g = p.grad
            if g is None:                                # unused in this phase (e.g. opacities once they are frozen)
                if id(p) not in self._zeros:
                    self._zeros[id(p)] = torch.zeros(p.numel(), dtype=self.flat.dtype, device=self.flat.device)
                g = self._zeros[id(p)]
            pieces.append(g.reshape(-1))
        with torch.no_grad():
            torch.cat(pieces, out=self.flat)
        for p, v in zip(self.params, self.views):
            p.grad = v

    def all_reduce(self, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)

    @property
    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()


class ViewParallel:
    """Wraps a model whose `forward(inp, labels)` returns a dict of losses incl. 'total' (the DBW contract,
    src/trainer.py:141-143) and which exposes `n_total_views` and `noise_generator` attributes."""

    VIEW_KEYS = ('imgs', 'R', 'T')

    def __init__(self, model, group=None, seed=227391):
        self.model, self.group, self.seed = model, group, seed
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bucket = GradBucket(model.parameters())
        self._step = 0

    def shard(self, inp):
        B = len(inp['imgs'])
        sl = shard_views(B, self.world_size, self.rank)
        out = dict(inp)
        for k in self.VIEW_KEYS:
            if k in inp:
                out[k] = inp[k][sl]
        return out, B

    def _sync_rng(self, device):
        g = getattr(self.model, 'noise_generator', None)
        if g is None or g.device != torch.device(device):
            g = torch.Generator(device=device)
            self.model.noise_generator = g
        g.manual_seed(self.seed + self._step)      # identical on every rank, fresh every step

    def forward_backward(self, inp, labels=None, already_sharded=False, n_total_views=None):
        """zero grads -> local forward -> backward -> ONE all-reduce.  Returns the (local) loss dict."""
        if not already_sharded:
            inp, n_total_views = self.shard(inp)
        self.model.n_total_views = n_total_views
        self._sync_rng(inp['imgs'].device)
        self._step += 1
        losses = self.model(inp, labels)
        self.bucket.backward(self.weighted_total(losses, len(inp['imgs']), n_total_views))
        self.bucket.all_reduce(self.group)
        return losses

    def weighted_total(self, losses, n_local_views, n_total_views):
        """the scalar each rank back-propagates: its share of the per-view terms + 1/world of the replicated terms"""
        shared = [v for k, v in losses.items() if k not in ('rgb', 'perceptual', 'total')]
        total = losses['rgb'] if 'rgb' in losses else 0.
        if 'perceptual' in losses:
            total = total + losses['perceptual'] * (n_local_views / float(n_total_views))
        if shared:
            total = total + sum(shared) / self.world_size
        return total

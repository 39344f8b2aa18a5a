"""Data-parallel gradient synchronisation (the ONE collective of the path: the reference's
tf.distribute.MirroredStrategy all-reduce, train_model.py:167-170, misc.py:27-58).

One process per GPU (torchrun), NCCL over NVLink/NVSwitch through torch.distributed. The flat fp32
gradient buffer of ParamTable is cut into contiguous buckets; a bucket's all-reduce is issued the
moment the LAST backward kernel that contributes to any of its parameters has been enqueued, so the
transfer overlaps the rest of the backward pass (torch's NCCL work runs on its own stream and only
waits for the compute stream's work enqueued so far). Every rank walks the same tape, hence all
ranks issue the buckets in the same order.

Loss scaling follows MirroredStrategy + Keras: each replica's loss terms are scaled 1/R (done in
M1.train_step through the gradient seeds), gradients are SUMMED across replicas."""
import torch
import torch.distributed as dist


class BucketedGradSync:
    def __init__(self, params, bucket_bytes=32 << 20, group=None):
        self.params = params
        self.group = group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        # buckets over the flat buffer in layout order; a parameter never straddles a bucket
        cap = max(1, bucket_bytes // 4)
        self.buckets = []          # [start, end)
        self.bucket_of = {}
        start = end = 0
        specs = sorted(params.specs.values(), key=lambda sp: sp.offset)
        for i, sp in enumerate(specs):
            nxt = specs[i + 1].offset if i + 1 < len(specs) else params.total
            if nxt - start > cap and end > start:
                self.buckets.append((start, end))
                start = end
            self.bucket_of[sp.name] = len(self.buckets)
            end = nxt
        if end > start:
            self.buckets.append((start, end))
        self.pending = []
        self.param_left = {}
        self.works = []
        self.fired = []
        self.flat = None
        self.pre_fire = None       # called before a bucket's all-reduce is issued (Engine.join_side: weight
                                   # gradients launched on the engine's side stream must have been joined)

    def begin(self, flat_grad, uses):
        """uses: {param name: number of backward closures that touch it this step}."""
        self.flat = flat_grad
        self.param_left = dict(uses)
        self.pending = [0] * len(self.buckets)
        for name, b in self.bucket_of.items():
            if self.param_left.get(name, 0) > 0:
                self.pending[b] += 1
        self.works, self.fired = [], []
        # buckets none of whose parameters receive a gradient this step still take part (zeros)

    def _fire(self, b):
        s, e = self.buckets[b]
        self.fired.append(b)
        if self.pre_fire is not None:
            self.pre_fire()
        if self.world_size > 1:
            self.works.append(dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, group=self.group,
                                              async_op=True))

    def param_done(self, name):
        left = self.param_left.get(name, 0)
        if left <= 0:
            return
        self.param_left[name] = left - 1
        if left == 1:
            b = self.bucket_of[name]
            self.pending[b] -= 1
            if self.pending[b] == 0:
                self._fire(b)

    def finish(self):
        for b in range(len(self.buckets)):
            if b not in self.fired:
                self._fire(b)
        for w in self.works:
            w.wait()
        self.works = []


def init_from_env(backend=None):
    """torchrun environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*) -> (rank, local_rank, world)."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, local, world

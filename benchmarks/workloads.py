"""Synthetic app_states of the BASELINE.json configurations, shared by both bench arms.

Every builder takes the snapshot module ``T`` it is built for — ``torchsnapshot_b200`` (this repo's mirror of the
reference interface) or the unmodified ``torchsnapshot`` staged under ``oracle/_ref`` — and only uses the names the
two have in common (``T.StateDict``), so that both arms checkpoint the very same objects.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch


# ---- C3: FSDP-layout Llama-3-8B bf16 (vocab 128256, dim 4096, 32 layers, ffn 14336, 8 KV heads) ---------------
def llama3_8b_shapes() -> List[Tuple[str, Tuple[int, ...]]]:
    dim, ffn, vocab, layers, kv = 4096, 14336, 128256, 32, 1024
    shapes = [("tok_embeddings.weight", (vocab, dim))]
    for i in range(layers):
        p = f"layers.{i}."
        shapes += [
            (p + "attention.wq.weight", (dim, dim)),
            (p + "attention.wk.weight", (kv, dim)),
            (p + "attention.wv.weight", (kv, dim)),
            (p + "attention.wo.weight", (dim, dim)),
            (p + "feed_forward.w1.weight", (ffn, dim)),
            (p + "feed_forward.w2.weight", (dim, ffn)),
            (p + "feed_forward.w3.weight", (ffn, dim)),
            (p + "attention_norm.weight", (dim,)),
            (p + "ffn_norm.weight", (dim,)),
        ]
    shapes += [("norm.weight", (dim,)), ("output.weight", (vocab, dim))]
    return shapes


def local_rows(rows: int, rank: int, world: int) -> Tuple[int, int]:
    """dim-0 chunk of ChunkShardingSpec / FSDP sharded state dicts: ceil split, ragged tail."""
    split = -(-rows // world)
    lo = min(rank * split, rows)
    return lo, max(0, min(split, rows - lo))


def build_llama_local(rank: int, world: int, device: torch.device, seed: int = 42, shapes=None):
    """{name: (local bf16 tensor, global shape, row offset)} — synthetic weights, random init (seed + rank)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + rank)
    out = {}
    for name, shape in shapes or llama3_8b_shapes():
        lo, n = local_rows(shape[0], rank, world)
        local = torch.empty((n,) + tuple(shape[1:]), dtype=torch.bfloat16, device=device)
        if local.numel():
            local.normal_(generator=gen)
        out[name] = (local, shape, lo)
    return out


def wrap_sharded(local, rank: int, device: torch.device):
    """FSDP1 SHARDED_STATE_DICT layout: one ShardedTensor per parameter, this rank's dim-0 slice as its only shard."""
    from torch.distributed._shard.sharded_tensor import Shard, ShardedTensor, ShardMetadata

    state = {}
    for name, (t, shape, lo) in local.items():
        off = [lo] + [0] * (len(shape) - 1)
        md = ShardMetadata(shard_offsets=off, shard_sizes=list(t.shape), placement=f"rank:{rank}/{device}")
        state[name] = ShardedTensor._init_from_local_shards([Shard(tensor=t, metadata=md)], tuple(shape))
    return state


def wrap_dtensor(local, world: int, device: torch.device):
    """FSDP2 layout: DTensor [Shard(0)] over a 1-D mesh."""
    from torch.distributed.device_mesh import init_device_mesh
    from torch.distributed.tensor import DTensor, Shard as ShardPlacement

    mesh = init_device_mesh(device.type, (world,))
    return {
        name: DTensor.from_local(t, mesh, [ShardPlacement(0)], run_check=False, shape=torch.Size(shape), stride=torch.empty(shape, device="meta").stride())
        for name, (t, shape, lo) in local.items()
    }


# ---- C2: DDP ResNet-50 + Adam, replicated -------------------------------------------------------------------
def resnet50_state_shapes() -> List[Tuple[str, Tuple[int, ...], torch.dtype]]:
    """state_dict of torchvision's resnet50 (25 557 032 parameters; 161 parameter tensors, 106 BN running stats,
    53 int64 num_batches_tracked scalars), written out so that the benchmark does not depend on torchvision."""
    out: List[Tuple[str, Tuple[int, ...], torch.dtype]] = []

    def conv(name, cout, cin, k):
        out.append((name + ".weight", (cout, cin, k, k), torch.float32))

    def bn(name, c):
        out.append((name + ".weight", (c,), torch.float32))
        out.append((name + ".bias", (c,), torch.float32))
        out.append((name + ".running_mean", (c,), torch.float32))
        out.append((name + ".running_var", (c,), torch.float32))
        out.append((name + ".num_batches_tracked", (), torch.int64))

    conv("conv1", 64, 3, 7)
    bn("bn1", 64)
    inplanes = 64
    for li, (planes, blocks) in enumerate(((64, 3), (128, 4), (256, 6), (512, 3)), start=1):
        for b in range(blocks):
            p = f"layer{li}.{b}"
            conv(p + ".conv1", planes, inplanes, 1)
            bn(p + ".bn1", planes)
            conv(p + ".conv2", planes, planes, 3)
            bn(p + ".bn2", planes)
            conv(p + ".conv3", planes * 4, planes, 1)
            bn(p + ".bn3", planes * 4)
            if b == 0:
                conv(p + ".downsample.0", planes * 4, inplanes, 1)
                bn(p + ".downsample.1", planes * 4)
            inplanes = planes * 4
    out.append(("fc.weight", (1000, 2048), torch.float32))
    out.append(("fc.bias", (1000,), torch.float32))
    return out


class ShapeListModule(torch.nn.Module):
    """A module whose state_dict has exactly the given (name, shape, dtype) list: float tensors become parameters,
    integer ones buffers.  Wrapped in DDP it makes ``_infer_replicated`` fire like a real model
    (T:snapshot.py:897-912, T:benchmarks/ddp/main.py:18-27,47-48)."""

    def __init__(self, shapes, device, seed: int = 42) -> None:
        super().__init__()
        gen = torch.Generator(device=device)
        gen.manual_seed(seed)  # identical on every rank: replicated state
        self._names = []
        for name, shape, dtype in shapes:
            key = name.replace(".", "__")
            if dtype.is_floating_point and not name.endswith(("running_mean", "running_var")):
                t = torch.empty(shape, dtype=dtype, device=device).normal_(generator=gen)
                self.register_parameter(key, torch.nn.Parameter(t))
            elif dtype.is_floating_point:
                self.register_buffer(key, torch.empty(shape, dtype=dtype, device=device).normal_(generator=gen))
            else:
                self.register_buffer(key, torch.full(shape, 1000, dtype=dtype, device=device))
            self._names.append(key)

    def forward(self, x):  # never trained here; DDP only needs a module with parameters
        return x


def build_c2(T, rank: int, world: int, device: torch.device, local_rank: int):
    """app_state {"model": DDP(resnet50-shaped module), "optim": Adam with exp_avg/exp_avg_sq/step filled in},
    replicated=["**"] (T:tests/test_ddp.py:74-78)."""
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP

    model = ShapeListModule(resnet50_state_shapes(), device)
    wrapped = DDP(model, device_ids=[local_rank]) if world > 1 and dist.is_initialized() and dist.get_backend() == "nccl" else model
    opt = torch.optim.Adam(wrapped.parameters(), lr=1e-3)
    gen = torch.Generator(device=device)
    gen.manual_seed(7)
    for p in wrapped.parameters():
        # what one optimizer.step() leaves behind (Adam keeps `step` as a 0-d fp32 CPU tensor by default)
        opt.state[p] = {
            "step": torch.tensor(1.0),
            "exp_avg": torch.empty_like(p).normal_(generator=gen),
            "exp_avg_sq": torch.empty_like(p).normal_(generator=gen).abs_(),
        }
    app_state = {"model": wrapped, "optim": opt}
    payload = 0
    for t in list(wrapped.state_dict().values()):
        payload += t.numel() * t.element_size()
    for st in opt.state.values():
        for v in st.values():
            payload += v.numel() * v.element_size()
    return app_state, {"replicated": ["**"]}, payload


# ---- C4: GPT-2-medium DDP training loop with async_take -------------------------------------------------------
class GPT2Medium(torch.nn.Module):
    """24 layers, d=1024, 16 heads, vocab 50257: 354.8 M fp32 parameters (tied output embedding)."""

    def __init__(self, vocab: int = 50257, d: int = 1024, layers: int = 24, ctx: int = 1024) -> None:
        super().__init__()
        self.vocab = vocab
        self.emb = torch.nn.Embedding(vocab, d)
        self.pos = torch.nn.Embedding(ctx, d)
        layer = torch.nn.TransformerEncoderLayer(d, 16, 4 * d, dropout=0.0, batch_first=True, norm_first=True)
        self.blocks = torch.nn.TransformerEncoder(layer, layers)
        self.ln = torch.nn.LayerNorm(d)

    def forward(self, x):
        h = self.emb(x) + self.pos(torch.arange(x.shape[1], device=x.device))
        return self.ln(self.blocks(h)) @ self.emb.weight.t()


# ---- C5: torchrec-style row-wise embedding table ------------------------------------------------------------
C5_ROWS, C5_COLS = 31_250_000, 128  # fp32: 16.0 GB


def c5_content(lo: int, n: int, cols: int, device) -> torch.Tensor:
    """Closed-form table content: every element can be checked on any rank after any resharding."""
    r = torch.arange(lo, lo + n, device=device, dtype=torch.int64).unsqueeze(1)
    c = torch.arange(cols, device=device, dtype=torch.int64).unsqueeze(0)
    return ((r * 131 + c * 7) % 65521).to(torch.float32)


def c5_sharded(t: torch.Tensor, lo: int, rows: int, cols: int, rank: int, device, process_group=None):
    from torch.distributed._shard.sharded_tensor import Shard, ShardedTensor, ShardMetadata

    md = ShardMetadata(shard_offsets=[lo, 0], shard_sizes=[t.shape[0], cols], placement=f"rank:{rank}/{device}")
    return ShardedTensor._init_from_local_shards([Shard(tensor=t, metadata=md)], (rows, cols), process_group=process_group)


def state_payload_bytes(state: Dict[str, torch.Tensor]) -> int:
    return sum(t.numel() * t.element_size() for t in state.values())

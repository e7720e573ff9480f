"""RAJNIViTWrapper — drop-in for rajni/wrapper/model.py:6-69, running on sm_100a kernels.

The class keeps the reference's constructor, attribute names (``m``, ``blocks``,
``pruning_schedule``, ``blk.attn`` swapped for ``RAJNIAttention``, ``blk.has_pruner``),
``forward`` and ``get_last_stats()``.  The body of ``forward`` does not call the base
model's modules: it reads their leaf parameters and launches the kernels of
``librajni_b200.so`` — pruned and un-pruned blocks, patch-embed and head alike.

Deliberate deviations from the reference:
  * schedule keys may be ints or digit strings (the reference silently prunes
    nothing for the string keys that ``json.load`` produces);
  * activations are bf16 with fp32 accumulation, scores are fp32;
  * ties in the top-k are broken by lower index (torch.topk leaves it undefined);
  * unsupported base models raise instead of silently diverging.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .. import ops
from ..packing import PackCache
from .attention import RAJNIAttention, keep_count


def _normalise_schedule(schedule: Dict) -> Dict[int, Dict]:
    out = {}
    for k, cfg in schedule.items():
        if isinstance(k, str):
            if not k.strip().lstrip("-").isdigit():
                raise ValueError(f"pruning_schedule key {k!r} is not a block index")
            k = int(k)
        if "keep_ratio" not in cfg:
            raise KeyError("keep_ratio")
        out[int(k)] = cfg
    return out


def _is_identity(m) -> bool:
    return m is None or isinstance(m, nn.Identity) or (isinstance(m, nn.Dropout) and m.p == 0.0)


AUTO_GRAPH_MAX_ROWS = 16384       # batch x tokens up to which forward() replays a CUDA graph by default


class RAJNIViTWrapper(nn.Module):
    def __init__(self, base_model: nn.Module, pruning_schedule: Dict[int, Dict]):
        super().__init__()
        self.m = base_model
        self.blocks = base_model.blocks
        self.pruning_schedule = pruning_schedule
        sched = _normalise_schedule(pruning_schedule)

        for i, blk in enumerate(self.blocks):
            if i in sched:
                cfg = sched[i]
                blk.attn = RAJNIAttention(blk.attn, keep_ratio=cfg["keep_ratio"], update=cfg.get("update", True))
                blk.has_pruner = True
            else:
                blk.has_pruner = False

        self._last_stats = None
        self._last_keep_idx: List[Optional[torch.Tensor]] = []
        self._packs = PackCache()
        self._ws = {}
        self._graphs = {}
        self._param_list = None
        # None = automatic: replay a captured CUDA graph when the batch is small enough to be launch-bound
        env = os.environ.get("RAJNI_CUDA_GRAPH", "")
        self.use_cuda_graph: Optional[bool] = None if env == "" else env != "0"
        self.input_norm: Optional[tuple] = None       # (mean[3], std[3]) for uint8 images, see set_input_normalization
        # Stream-K tail of the fc2 GEMMs (gemm_tcgen05.cu): +1.4 % at 32 images per GPU, nothing at 64 and above.  OFF by default
        # because the fp32 summation order of a split tile depends on the tile count, i.e. on the batch size: without it an
        # image's logits are bit-identical whatever batch it is evaluated in (tests/test_gpu_e2e.py::test_full_batch_properties).
        self.stream_k: bool = os.environ.get("RAJNI_STREAM_K", "0") == "1"
        self._validate()

    def set_input_normalization(self, mean, std) -> "RAJNIViTWrapper":
        """Extension (not in the reference): accept raw uint8 pixels [B,3,S,S] and apply the loader's ToTensor + Normalize
        (run.py:62-70) inside the patch kernel - a quarter of the host-to-device bytes of fp32 images, same logits."""
        mean, std = tuple(float(v) for v in mean), tuple(float(v) for v in std)
        if len(mean) != 3 or len(std) != 3 or min(std) <= 0:
            raise ValueError("mean and std must have three entries, std positive")
        self.input_norm = (mean, std)
        self._graphs = {}
        return self

    # ------------------------------------------------------------------ contract checks
    def _validate(self):
        m = self.m
        proj = getattr(m.patch_embed, "proj", None)
        if not isinstance(proj, nn.Conv2d) or proj.kernel_size != (16, 16) or proj.stride != (16, 16) or proj.in_channels != 3:
            raise NotImplementedError("patch_embed.proj must be Conv2d(3, C, kernel=16, stride=16)")
        if not _is_identity(getattr(m.patch_embed, "norm", None)):
            raise NotImplementedError("patch_embed.norm is not supported")
        for name in ("norm_pre", "fc_norm", "patch_drop"):
            if not _is_identity(getattr(m, name, None)):
                raise NotImplementedError(f"{name} = {type(getattr(m, name)).__name__} is not supported")
        if getattr(m, "global_pool", "token") != "token":
            raise NotImplementedError("only global_pool='token' (CLS) is supported")
        if not isinstance(m.norm, nn.LayerNorm) or not isinstance(m.head, nn.Linear):
            raise NotImplementedError("norm must be LayerNorm and head must be Linear")
        C = proj.out_channels
        for i, blk in enumerate(self.blocks):
            for name in ("ls1", "ls2", "drop_path1", "drop_path2"):
                if not _is_identity(getattr(blk, name, None)):
                    raise NotImplementedError(f"blocks[{i}].{name} = {type(getattr(blk, name)).__name__} is not supported")
            attn = blk.attn
            if attn.num_heads * 64 != C:
                raise NotImplementedError(f"blocks[{i}]: head dim must be 64 (C={C}, heads={attn.num_heads})")
            if abs(float(attn.scale) - 0.125) > 1e-12:
                raise NotImplementedError(f"blocks[{i}]: attention scale {attn.scale} != 64**-0.5")
            for name in ("q_norm", "k_norm"):
                if not _is_identity(getattr(attn, name, None)):
                    raise NotImplementedError(f"blocks[{i}].attn.{name} is not supported")
            act = blk.mlp.act
            if not isinstance(act, nn.GELU) or getattr(act, "approximate", "none") != "none":
                raise NotImplementedError(f"blocks[{i}].mlp.act must be exact (erf) nn.GELU")
            if not _is_identity(getattr(blk.mlp, "norm", None)):
                raise NotImplementedError(f"blocks[{i}].mlp.norm is not supported")
            if not isinstance(blk.norm1, nn.LayerNorm) or not isinstance(blk.norm2, nn.LayerNorm):
                raise NotImplementedError(f"blocks[{i}]: norm1/norm2 must be LayerNorm")

    def _apply(self, fn, *a, **kw):
        self._packs.clear()
        self._ws.clear()
        self._graphs = {}
        self._param_list = None
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, *a, **kw):
        self._graphs = {}
        self._param_list = None
        return super().load_state_dict(*a, **kw)

    def get_last_stats(self):
        return self._last_stats

    # ------------------------------------------------------------------ workspace
    def _workspace(self, B: int, S: int, dev: torch.device):
        key = (B, S, dev, self.stream_k)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        m = self.m
        C = m.patch_embed.proj.out_channels
        G = S // 16
        P = G * G
        N0 = P + 1
        if m.pos_embed.shape[1] < N0:
            raise ValueError(f"pos_embed has {m.pos_embed.shape[1]} positions, input needs {N0}")
        hidden = max(blk.mlp.fc1.out_features for blk in self.blocks)
        bf = dict(device=dev, dtype=torch.bfloat16)
        idx = torch.arange(B * P, device=dev, dtype=torch.int32)
        ws = dict(
            P=P, N0=N0, C=C,
            cols=torch.empty((B * P, 768), **bf),
            xa=torch.empty((B * N0, C), **bf), xb=torch.empty((B * N0, C), **bf),
            qkv=torch.empty((B * N0, 3 * C), **bf),
            att=torch.empty((B * N0, C), **bf), hid=torch.empty((B * N0, hidden), **bf),
            cls=torch.empty((B, C), **bf),
            embed_out_map=((idx // P) * N0 + 1 + idx % P).to(torch.int32),
            embed_pos_map=(1 + idx % P).to(torch.int32),
            # LayerNorm partial sums (sum, sum of squares) of the rows of x, one slot per (n-tile, column half)
            # of the GEMM that produced them; consumed by the LN-folded qkv / fc1 GEMMs
            stat_slots=ops.row_stats_slots(C),
            stats=torch.zeros((ops.row_stats_slots(C), B * N0, 2), device=dev, dtype=torch.float32),
            sel={},
            # stream-K scratch of the long-K GEMMs (fc2), only when asked for: owned here like every other pointer a captured
            # graph bakes in
            gemm_ws=ops.gemm_workspace(dev) if self.stream_k else None,
            # scratch of the split score path (small batches): owned here, so a captured graph's pointer lives with the graph
            score_ws=(torch.empty(max(ops.score_workspace_bytes(B, N0, C, blk.attn.num_heads) for blk in self.blocks),
                                  device=dev, dtype=torch.uint8) if B <= ops.SPLIT_SCORE_MAX_BATCH else None),
        )
        self._ws = {key: ws}        # keep one shape resident
        self._graphs = {}           # a captured graph points into the workspace it was captured with
        return ws

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """model.py:30-69.  The launch sequence of a given input shape can be captured once into a CUDA graph and replayed:
        the 68 launches of a step cost ~13 us of host time each, which dominates small batches (vit_tiny at batch 8:
        0.88 -> 0.53 ms per step).  ``use_cuda_graph`` = None (default) replays when the batch has at most
        ``AUTO_GRAPH_MAX_ROWS`` token rows, True / False force it (RAJNI_CUDA_GRAPH=1 / 0 in the environment do the same).
        A captured graph is dropped when any parameter changes (storage or version), on ``.to()`` / ``.bfloat16()`` /
        ``load_state_dict`` and when the input normalisation changes; results are bit-identical to the eager sequence."""
        if x.device.type == "cuda" and x.device.index is not None and x.device.index != torch.cuda.current_device():
            with torch.cuda.device(x.device):          # the C ABI launches on the current device
                return self.forward(x)
        use = self.use_cuda_graph
        if use is None and x.dim() == 4:
            use = x.shape[0] * ((x.shape[2] // 16) * (x.shape[3] // 16) + 1) <= AUTO_GRAPH_MAX_ROWS
        if not use or x.device.type != "cuda" or x.dim() != 4 or ops._prof is not None:      # (the per-kernel profiler times eager launches)
            return self._forward_eager(x)
        # what _forward_eager reads besides the input: the schedule of the pruned blocks (attention.py:25-32) ...
        sched = tuple((blk.attn.keep_ratio, blk.attn.update) if blk.has_pruner else None for blk in self.blocks)
        key = (tuple(x.shape), x.dtype, x.device, self.training, sched, self.input_norm, self.stream_k)
        # ... and the parameters.  Storage changes go through _apply / load_state_dict (which drop the graphs); in-place
        # updates bump the tensors' version counters, summed here over a parameter list that is built once.
        if self._param_list is None:
            self._param_list = list(self.parameters())
        wsig = sum(q._version for q in self._param_list)
        entry = self._graphs.get(key)
        if entry is None or entry[5] != wsig:
            x_static = x.clone()
            self._forward_eager(x_static)                       # builds packs and workspace, warms every kernel
            torch.cuda.synchronize(x.device)
            graph = torch.cuda.CUDAGraph()
            # thread-local capture mode: a DataLoader's pin-memory thread may call into CUDA while this thread captures
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                y_static = self._forward_eager(x_static)
            entry = (graph, x_static, y_static, self._last_stats, self._last_keep_idx, wsig)
            self._graphs = {key: entry}                         # one shape resident, like the workspace
        graph, x_static, y_static, stats, keep, _ = entry
        x_static.copy_(x)
        graph.replay()
        self._last_stats, self._last_keep_idx = stats, keep
        return y_static.clone()

    @torch.no_grad()
    def _forward_eager(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != x.shape[3] or x.shape[2] % 16:
            raise ValueError(f"expected images [B,3,S,S] with S a multiple of 16, got {tuple(x.shape)}")
        if x.device.type != "cuda":
            raise RuntimeError("RAJNIViTWrapper (B200) runs on CUDA tensors only; there is no CPU path")
        if self.m.pos_embed.device != x.device or self.m.head.weight.device != x.device:
            raise RuntimeError(f"Expected all tensors to be on the same device: input on {x.device}, model on "
                               f"{self.m.pos_embed.device} (call .to(device) on the wrapper first)")
        if self.training:
            for blk in self.blocks:
                for d in (blk.attn.proj_drop, getattr(blk.mlp, "drop1", None), getattr(blk.mlp, "drop2", None)):
                    if isinstance(d, nn.Dropout) and d.p > 0:
                        raise NotImplementedError("dropout in training mode is not supported (inference path)")
        m = self.m
        out_dtype = x.dtype if x.dtype.is_floating_point else torch.float32
        if x.dtype == torch.uint8:
            if self.input_norm is None:
                raise TypeError("uint8 images need set_input_normalization(mean, std) first (the reference takes normalised floats)")
        elif x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        x = x.contiguous()
        B, _, S, _ = x.shape
        ws = self._workspace(B, S, x.device)
        P, N, C = ws["P"], ws["N0"], ws["C"]
        pk = self._packs

        # ---- patch + CLS + pos (model.py:31-37): im2col, then one GEMM whose epilogue adds
        #      bias and pos_embed and scatters into rows 1..P of each image
        pe_w, pe_b = pk.linear(m.patch_embed.proj)
        def _pos_pack():
            pos_ = m.pos_embed.detach()[0, : P + 1].to(torch.bfloat16).contiguous()
            cls_ = (m.cls_token.detach()[0, 0] + m.pos_embed.detach()[0, 0]).to(torch.bfloat16).contiguous()
            c32 = cls_.to(torch.float32)
            return pos_, cls_, float(c32.sum()), float((c32 * c32).sum())

        pos, cls_pos0, cls_sum, cls_sumsq = pk.tensors(("pos", P), (m.pos_embed, m.cls_token), _pos_pack)
        cur, nxt = ws["xa"], ws["xb"]
        stats, slots = ws["stats"], ws["stat_slots"]
        ops.patch_im2col(x, 16, ws["cols"], cls_pos0, cur, C, row_stats=stats, stats_slots=slots,
                         cls_sum=cls_sum, cls_sumsq=cls_sumsq, norm=self.input_norm if x.dtype == torch.uint8 else None)
        ops.gemm(ws["cols"], pe_w, pe_b, B * P, C, 768, residual=pos, ldres=C, res_row_map=ws["embed_pos_map"],
                 out=cur, ldd=C, out_row_map=ws["embed_out_map"], row_stats=stats, tag="embed")

        # Zig-zag traversal: every persistent kernel walks its row tiles in the direction opposite to the kernel that
        # produced its main operand, so it starts on the rows that kernel wrote last and that still sit in L2.
        zig = os.environ.get("RAJNI_NO_ZIGZAG") is None       # A/B switch for profiling
        rev = zig                      # the embed GEMM walked forward
        scores = None
        token_counts = []
        keep_log: List[Optional[torch.Tensor]] = []
        for i, blk in enumerate(self.blocks):
            token_counts.append(N)                                                   # model.py:43
            H = blk.attn.num_heads
            # norm1 / norm2 are folded into the qkv / fc1 GEMMs: x itself is the A operand, the row statistics
            # come from the epilogue of whichever GEMM stored x (embed, proj or fc2)
            qw, qb, qsum, e1 = pk.linear_ln(blk.attn.qkv, blk.norm1)
            pw, pb = pk.linear(blk.attn.proj)
            f1w, f1b, f1sum, e2 = pk.linear_ln(blk.mlp.fc1, blk.norm2)
            f2w, f2b = pk.linear(blk.mlp.fc2)
            hidden = f1w.shape[0]
            M = B * N
            ops.gemm(cur, qw, qb, M, 3 * C, C, out=ws["qkv"], ln=(stats, slots, qsum, e1), tag="qkv", reverse=rev)   # model.py:51 + attention.py:22
            if blk.has_pruner:
                attn: RAJNIAttention = blk.attn
                keep = keep_count(N, attn.keep_ratio)                                 # attention.py:31-32
                if keep > N - 1:
                    raise RuntimeError("selected index k out of range")               # attention.py:35
                Np = keep + 1
                sel = ws["sel"].get(i)
                if sel is None or sel[0].shape[1] != Np:
                    sel = (torch.empty((B, Np), device=x.device, dtype=torch.int32),
                           torch.empty((B, Np), device=x.device, dtype=torch.float32),
                           torch.empty((B * Np,), device=x.device, dtype=torch.int32))
                    ws["sel"][i] = sel
                keep_idx, next_scores, row_map = sel
                if attn.update or scores is None:                                     # attention.py:25-28
                    ops.score_select(ws["qkv"][: B * N].view(B, N, 3 * C), H, keep,
                                     keep_idx=keep_idx, next_scores=next_scores, row_map=row_map, workspace=ws["score_ws"])
                else:
                    ops.select(scores, keep, keep_idx=keep_idx, next_scores=next_scores, row_map=row_map)
                if Np > ops.LONG_SEQ and "qkvc" not in ws:                             # compaction buffer, long sequences only
                    ws["qkvc"] = torch.empty((B * ws["N0"], 3 * C), device=x.device, dtype=torch.bfloat16)
                ops.attention(ws["qkv"], row_map, B, N, Np, C, H, float(attn.scale), out=ws["att"], reverse=zig and not rev,
                              compact=ws.get("qkvc"))
                # proj + gathered residual: x_new[b,j] = x[b, keep_idx[b,j]] + proj(att)   model.py:55-58
                ops.gemm(ws["att"], pw, pb, B * Np, C, C, residual=cur, ldres=C, res_row_map=row_map, out=nxt, ldd=C,
                         row_stats=stats, tag="proj", reverse=rev)
                cur, nxt = nxt, cur
                scores = next_scores                                                  # attention.py:58
                keep_log.append(keep_idx)
                N = Np
                M = B * N
            else:
                ops.attention(ws["qkv"], None, B, N, N, C, H, float(blk.attn.scale), out=ws["att"], reverse=zig and not rev)
                ops.gemm(ws["att"], pw, pb, M, C, C, residual=cur, ldres=C, out=cur, ldd=C, row_stats=stats, tag="proj", reverse=rev)   # in place
                scores = None                                                         # model.py:63
                keep_log.append(None)
            ops.gemm(cur, f1w, f1b, M, hidden, C, gelu=True, out=ws["hid"], ldd=hidden,
                     ln=(stats, slots, f1sum, e2), tag="fc1", reverse=zig and not rev)                                    # model.py:59 (norm2 + fc1 + GELU)
            ops.gemm(ws["hid"], f2w, f2b, M, C, hidden, residual=cur, ldres=C, out=cur, ldd=C, row_stats=stats, tag="fc2", reverse=rev,
                     workspace=ws["gemm_ws"] if self.stream_k else None)
            rev = zig and not rev

        # ---- final norm on the CLS rows only (LayerNorm is row-wise) + head   model.py:65-66
        gn, bn, en = pk.norm(m.norm)
        hw, hb = pk.linear(m.head)
        ops.layernorm(cur, gn, bn, en, B, C, in_row_stride=N * C, out=ws["cls"])
        logits = ops.gemm(ws["cls"], hw, hb, B, hw.shape[0], C, out_f32=True, tag="head")

        self._last_stats = {"token_counts": token_counts}                            # model.py:68
        self._last_keep_idx = keep_log
        return logits if out_dtype == torch.float32 else logits.to(out_dtype)

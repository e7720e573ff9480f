"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Never import this from the product package.

A functional restatement (plain torch on CPU, fp32 or fp64) of the reference's
token-pruning forward path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only
as the checker or as the timed CPU baseline.

Parity status: the reference ships no tests, fixtures or golden vectors
(SURVEY.md section 4), so parity is *unpinned by the reference's own tests*.  This
oracle is pinned instead against outputs of the unmodified reference imported
from /root/reference in the build container: ``tests/make_golden.py`` generated
``tests/golden/*.npz`` from it, and ``tests/test_oracle.py`` checks the oracle
against those files (and, when /root/reference is present, against the live
reference).

Each function cites the reference lines it restates.  The model is passed as a
flat ``dict`` of tensors (see ``extract_params``), not as an ``nn.Module``.
"""
from __future__ import annotations

import math
import time
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- score
def importance(qkv: Tensor, num_heads: int, eps: float = 1e-6) -> Tensor:
    """Per-token score, [B,N,3C] -> [B,N].   rajni/wrapper/importance.py:5-34

    score = mean_h softmax_n(q_cls . k / sqrt(D)) * sigmoid(zscore_n(|mean_h v - mean_n mean_h v|))
    The z-score uses the unbiased std (torch default) plus eps (importance.py:29).
    """
    B, N, C3 = qkv.shape
    C = C3 // 3
    D = C // num_heads
    planes = qkv.reshape(B, N, 3, num_heads, D)
    q_cls = planes[:, 0, 0]                         # [B,H,D]      importance.py:18
    k = planes[:, :, 1]                             # [B,N,H,D]
    v = planes[:, :, 2]                             # [B,N,H,D]
    logits = torch.einsum("bhd,bnhd->bhn", q_cls, k) / math.sqrt(D)   # importance.py:19
    a_cls = logits.softmax(dim=-1).mean(dim=1)      # [B,N]        importance.py:20-21
    vm = v.mean(dim=2)                              # [B,N,D]      importance.py:24
    vm = vm - vm.mean(dim=1, keepdim=True)          #              importance.py:25
    r = vm.norm(dim=-1)                             # [B,N]        importance.py:27
    mu = r.mean(dim=1, keepdim=True)                #              importance.py:28
    sd = r.std(dim=1, keepdim=True) + eps           #              importance.py:29
    return a_cls * torch.sigmoid((r - mu) / sd)     #              importance.py:31-34


def keep_count(num_tokens: int, keep_ratio: float) -> int:
    """Patches kept (CLS excluded).   rajni/wrapper/attention.py:31-32

    Python double multiply then truncation, e.g. int(0.72*120) == 86.
    """
    return max(1, int(keep_ratio * (num_tokens - 1)))


def select(scores: Tensor, keep: int) -> Tensor:
    """CLS-preserving ascending kept-token index, [B,N] -> int64 [B,keep+1].

    rajni/wrapper/attention.py:34-39 (topk over scores[:,1:], sort, +1, prepend 0).
    torch.topk's order among equal scores is implementation-defined; this oracle
    fixes the rule "greater score first, then lower index", realised with a stable
    descending sort.  It equals the reference whenever no tie straddles the cut.
    """
    B, N = scores.shape
    if keep > N - 1:
        raise RuntimeError("selected index k out of range")       # what topk raises, attention.py:35
    patch = scores[:, 1:]
    order = torch.sort(patch, dim=1, descending=True, stable=True).indices[:, :keep]
    idx = torch.sort(order, dim=1).values + 1
    cls = torch.zeros((B, 1), dtype=torch.long)
    return torch.cat([cls, idx], dim=1)


def tie_straddles_cut(scores: Tensor, keep: int) -> Tensor:
    """bool [B]: True where the keep-th and (keep+1)-th largest patch scores are equal,
    i.e. where the reference's own kept set is implementation-defined."""
    patch = scores[:, 1:]
    if keep >= patch.shape[1]:
        return torch.zeros(scores.shape[0], dtype=torch.bool)
    s = torch.sort(patch, dim=1, descending=True).values
    return s[:, keep - 1] == s[:, keep]


# ----------------------------------------------------------------------- attention
def mha(qkv: Tensor, num_heads: int, scale: float) -> Tensor:
    """Dense multi-head attention on a packed qkv [B,N,3C] -> [B,N,C].  attention.py:45-54"""
    B, N, C3 = qkv.shape
    C = C3 // 3
    q, k, v = qkv.reshape(B, N, 3, num_heads, C // num_heads).permute(2, 0, 3, 1, 4)
    p = ((q @ k.transpose(-2, -1)) * scale).softmax(dim=-1)
    return (p @ v).transpose(1, 2).reshape(B, N, C)


def pruned_attention(x_norm: Tensor, prev_scores: Optional[Tensor], blk: Dict[str, Tensor],
                     num_heads: int, keep_ratio: float, update: bool,
                     forced_idx: Optional[Tensor] = None
                     ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """RAJNIAttention.forward.   rajni/wrapper/attention.py:17-60

    Returns (out [B,Np,C], keep_idx [B,Np] int64, next_scores [B,Np], scores [B,N]).
    ``forced_idx`` (tests only) replaces the selection with a given kept-token index so the
    arithmetic downstream of the cut can be compared without selection noise; the oracle's own
    scores are still computed and returned.
    """
    B, N, C = x_norm.shape
    qkv = F.linear(x_norm, blk["qkv_w"], blk["qkv_b"])                  # attention.py:22
    if update or prev_scores is None:                                    # attention.py:25-28
        scores = importance(qkv, num_heads)
    else:
        scores = prev_scores
    keep = keep_count(N, keep_ratio)
    keep_idx = select(scores, keep) if forced_idx is None else forced_idx.to(torch.long)
    assert keep_idx.shape == (B, keep + 1)
    kept = torch.gather(qkv, 1, keep_idx[:, :, None].expand(-1, -1, 3 * C))   # attention.py:42-43
    scale = (C // num_heads) ** -0.5
    out = F.linear(mha(kept, num_heads, scale), blk["proj_w"], blk["proj_b"])   # attention.py:55
    next_scores = torch.gather(scores, 1, keep_idx)                      # attention.py:58
    return out, keep_idx, next_scores, scores


# --------------------------------------------------------------------------- model
def extract_params(model) -> Dict:
    """Flatten a timm-style ViT into plain tensors (the attribute surface of SURVEY 8b)."""
    def lin(m):
        return m.weight.detach(), (m.bias.detach() if m.bias is not None else None)

    blocks = []
    for blk in model.blocks:
        attn = blk.attn
        qw, qb = lin(attn.qkv)
        pw, pb = lin(attn.proj)
        f1w, f1b = lin(blk.mlp.fc1)
        f2w, f2b = lin(blk.mlp.fc2)
        blocks.append(dict(
            n1_w=blk.norm1.weight.detach(), n1_b=blk.norm1.bias.detach(), n1_eps=blk.norm1.eps,
            n2_w=blk.norm2.weight.detach(), n2_b=blk.norm2.bias.detach(), n2_eps=blk.norm2.eps,
            qkv_w=qw, qkv_b=qb, proj_w=pw, proj_b=pb,
            fc1_w=f1w, fc1_b=f1b, fc2_w=f2w, fc2_b=f2b, num_heads=attn.num_heads))
    pe_w, pe_b = lin(model.patch_embed.proj)
    hw, hb = lin(model.head)
    return dict(pe_w=pe_w, pe_b=pe_b, patch=model.patch_embed.proj.kernel_size[0],
                cls=model.cls_token.detach(), pos=model.pos_embed.detach(),
                norm_w=model.norm.weight.detach(), norm_b=model.norm.bias.detach(),
                norm_eps=model.norm.eps, head_w=hw, head_b=hb, blocks=blocks)


def cast_params(params: Dict, dtype) -> Dict:
    def c(v):
        return v.to(dtype) if isinstance(v, torch.Tensor) else v
    out = {k: c(v) for k, v in params.items() if k != "blocks"}
    out["blocks"] = [{k: c(v) for k, v in b.items()} for b in params["blocks"]]
    return out


def normalise_schedule(schedule: Dict) -> Dict[int, Tuple[float, bool]]:
    """{block: {"keep_ratio", "update"(=True)}} -> {int block: (ratio, update)}.  model.py:14-20

    Int keys only, exactly like the reference (string keys from json.load never match,
    SURVEY 4.3); callers wanting JSON schedules convert keys first.
    """
    return {i: (cfg["keep_ratio"], cfg.get("update", True))
            for i, cfg in schedule.items() if isinstance(i, int)}


def embed(params: Dict, images: Tensor) -> Tensor:
    """Patch-embed + CLS + position.   rajni/wrapper/model.py:31-37"""
    p = params["patch"]
    x = F.conv2d(images, params["pe_w"], params["pe_b"], stride=p).flatten(2).transpose(1, 2)
    x = torch.cat([params["cls"].expand(x.shape[0], -1, -1), x], dim=1)
    return x + params["pos"][:, : x.shape[1]]


def dense_block(x: Tensor, blk: Dict) -> Tensor:
    """Un-pruned timm block.   model.py:61-63"""
    C = x.shape[-1]
    H = blk["num_heads"]
    h = F.layer_norm(x, (C,), blk["n1_w"], blk["n1_b"], blk["n1_eps"])
    qkv = F.linear(h, blk["qkv_w"], blk["qkv_b"])
    x = x + F.linear(mha(qkv, H, (C // H) ** -0.5), blk["proj_w"], blk["proj_b"])
    return x + mlp(x, blk)


def mlp(x: Tensor, blk: Dict) -> Tensor:
    C = x.shape[-1]
    h = F.layer_norm(x, (C,), blk["n2_w"], blk["n2_b"], blk["n2_eps"])
    return F.linear(F.gelu(F.linear(h, blk["fc1_w"], blk["fc1_b"])), blk["fc2_w"], blk["fc2_b"])


def pruned_block(x: Tensor, scores: Optional[Tensor], blk: Dict, keep_ratio: float, update: bool,
                 forced_idx: Optional[Tensor] = None):
    """Pruned block: LN1 -> pruned attention -> gather residual -> + -> MLP -> +.  model.py:50-59"""
    C = x.shape[-1]
    h = F.layer_norm(x, (C,), blk["n1_w"], blk["n1_b"], blk["n1_eps"])
    out, keep_idx, next_scores, full_scores = pruned_attention(
        h, scores, blk, blk["num_heads"], keep_ratio, update, forced_idx)
    x = torch.gather(x, 1, keep_idx[:, :, None].expand(-1, -1, C)) + out
    x = x + mlp(x, blk)
    return x, keep_idx, next_scores, full_scores


@torch.no_grad()
def forward(params: Dict, images: Tensor, schedule: Dict, trace: Optional[List] = None,
            forced_keep: Optional[List[Optional[Tensor]]] = None) -> Tuple[Tensor, Dict]:
    """RAJNIViTWrapper.forward.   rajni/wrapper/model.py:30-69

    Returns (logits [B,classes], {"token_counts": [...]}); if ``trace`` is a list it
    receives one dict per block (block input, keep_idx, scores, output) for
    teacher-forced comparisons.  ``forced_keep`` (one entry per block, None on un-pruned blocks)
    teacher-forces the selection, see ``pruned_attention``.
    """
    sched = normalise_schedule(schedule)
    x = embed(params, images)
    scores = None
    counts = []
    for i, blk in enumerate(params["blocks"]):
        counts.append(x.shape[1])
        rec = {"block": i, "x_in": x} if trace is not None else None
        if i in sched:
            ratio, update = sched[i]
            prev = scores
            forced = forced_keep[i] if forced_keep is not None else None
            x, keep_idx, scores, full = pruned_block(x, scores, blk, ratio, update, forced)
            if rec is not None:
                rec.update(pruned=True, keep_idx=keep_idx, scores=full, prev_scores=prev,
                           next_scores=scores)
        else:
            x = dense_block(x, blk)
            scores = None                                        # model.py:63
            if rec is not None:
                rec.update(pruned=False)
        if rec is not None:
            rec["x_out"] = x
            trace.append(rec)
    C = x.shape[-1]
    cls = F.layer_norm(x[:, 0], (C,), params["norm_w"], params["norm_b"], params["norm_eps"])
    logits = F.linear(cls, params["head_w"], params["head_b"])  # model.py:65-66 (LN is row-wise)
    return logits, {"token_counts": counts}


@torch.no_grad()
def evaluate(params: Dict, schedule: Dict, batches, max_batches=None, warmup=5):
    """evaluate_model on the CPU.   rajni/eval.py:6-75  -> (acc %, images/s)"""
    batches = list(batches)
    for i in range(warmup):
        forward(params, batches[i % len(batches)][0], schedule)
    correct = total = 0
    elapsed = 0.0
    for i, (images, labels) in enumerate(batches):
        if max_batches is not None and i >= max_batches:
            break
        t0 = time.time()
        logits, _ = forward(params, images, schedule)
        elapsed += time.time() - t0
        correct += int((logits.argmax(dim=1) == labels).sum())
        total += labels.numel()
    return 100.0 * correct / max(total, 1), total / max(elapsed, 1e-6)

"""Architecture strings -> layer tables for the HVAE (same grammar as the reference's
``enc_arch`` / ``dec_arch`` flags, src/vae.py:90-120,198-218; presets in src/hps.py:12-78)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional


@dataclass
class EncStage:
    cin: int
    cmid: int
    cout: int
    down: Optional[int]
    res_in: int      # spatial size the block's convs run at
    res_out: int     # after pooling and odd-size zero padding


@dataclass
class DecStage:
    idx: int
    res: int
    cin: int
    cout: int
    cmid: int
    ksize: int
    stochastic: bool


def _tokens(arch: str):
    for tok in arch.split(","):
        res, body = tok.split("b", 1)
        yield int(res), body


def encoder_plan(args) -> List[EncStage]:
    widths = list(args.widths)
    seq = []  # (out width, down rate)
    for si, (_, body) in enumerate(_tokens(args.enc_arch)):
        n, _, rate = body.partition("d")
        if si == 0 and int(n) == 0 and not rate:
            raise NotImplementedError("stride-2 stem variant is dead code in the reference (SURVEY Q4)")
        seq += [(widths[si], None)] * int(n)
        if rate:
            seq.append((widths[si + 1], int(rate[0])))
    plan: List[EncStage] = []
    res = int(args.input_res)
    for i, (w, d) in enumerate(seq):
        w_in = seq[max(i - 1, 0)][0]
        r_out = res // d if d else res
        if r_out % 2 == 1 and r_out > 1:
            r_out += 1  # src/vae.py:130-132
        plan.append(EncStage(w_in, int(w_in / args.bottleneck), w, d, res, r_out))
        res = r_out
    return plan


def decoder_plan(args) -> List[DecStage]:
    rev = list(args.widths)[::-1]
    seq = []
    for si, (res, body) in enumerate(_tokens(args.dec_arch)):
        seq += [(res, rev[si])] * int(body)
    plan: List[DecStage] = []
    for i, (res, w) in enumerate(seq):
        w_next = seq[min(i + 1, len(seq) - 1)][1]
        plan.append(DecStage(i, res, w, w_next, int(w / args.bottleneck), 3 if res > 2 else 1,
                             res <= args.z_max_res))
    return plan


def bias_resolutions(args):
    """(resolution, width) of the learned per-resolution biases, src/vae.py:211-218"""
    rev = list(args.widths)[::-1]
    all_res = sorted({res for res, _ in _tokens(args.dec_arch)})
    return [(r, rev[i]) for i, r in enumerate(all_res) if r <= args.bias_max_res]

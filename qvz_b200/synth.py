"""Seeded synthetic quality-value files (SURVEY.md section 8d / BASELINE.md section 3).

Fixed-length lines of C bytes '!'+q, q in [0, 41], '\\n'-terminated; per read a latent class with a
well separated mean (so k-means has structure and no cluster runs empty), a position-dependent target
decaying along the read, an AR(1) walk around it and ~2 % drop-outs to Q2.  torch is used only as an
array library here (same code on CPU for the tests and on the GPU for bench.py).
"""
from __future__ import annotations

import torch

CLASS_WEIGHTS = (0.40, 0.25, 0.15, 0.12, 0.08)
CLASS_OFFSET = 6.0
CHUNK = 1_000_000

# BASELINE.json configs -> (lines, columns, clusters, profile)
CONFIGS = {
    "cfg1": dict(lines=1_000_000, columns=100, clusters=1, profile="illumina", mode="ratio", ratio=1.0, dist="M"),
    "cfg2": dict(lines=20_000_000, columns=150, clusters=1, profile="illumina", mode="fixed", ratio=2.0, dist="L"),
    "cfg3": dict(lines=50_000_000, columns=150, clusters=3, profile="illumina", mode="ratio", ratio=0.5, dist="A", threshold=4),
    "cfg4": dict(lines=200_000_000, columns=150, clusters=5, profile="illumina", mode="ratio", ratio=1.0, dist="M"),
    "cfg5": dict(lines=40_000_000, columns=250, clusters=2, profile="miseq", mode="fixed", ratio=4.0, dist="M"),
}


def _targets(columns: int, profile: str, device) -> torch.Tensor:
    pos = torch.arange(columns, dtype=torch.float32, device=device) / max(columns - 1, 1)
    if profile == "miseq":                       # flat head, steep tail decay to ~15
        return torch.where(pos < 0.6, torch.full_like(pos, 38.0), 38.0 - 23.0 * ((pos - 0.6) / 0.4) ** 1.5)
    return 40.0 - 15.0 * pos                     # Illumina-like: ~40 -> ~25


def _chunk(n: int, columns: int, seed: int, ci: int, tgt, w, device, blk: torch.Tensor) -> None:
    """Lines [ci*CHUNK, ci*CHUNK + n) of the file with this seed -> blk [n, columns+1]."""
    g = torch.Generator(device=device)
    g.manual_seed(seed * 1_000_003 + ci)
    cls = torch.multinomial(w, n, replacement=True, generator=g).to(torch.float32) * CLASS_OFFSET
    q = None
    for c in range(columns):
        t = tgt[c] - cls
        noise = torch.randn(n, generator=g, device=device) * 2.0
        q = (t + noise) if q is None else (0.7 * q + 0.3 * t + noise)
        q = q.round_().clamp_(0.0, 41.0)
        drop = torch.rand(n, generator=g, device=device) < 0.02
        blk[:, c] = torch.where(drop, torch.full_like(q, 2.0), q).to(torch.uint8) + 33
    blk[:, columns] = 10


def synth_rows(lines: int, columns: int, seed: int = 1234, profile: str = "illumina",
               device: str | torch.device = "cpu", out: torch.Tensor | None = None,
               first_line: int = 0, total_lines: int | None = None) -> torch.Tensor:
    """uint8 [lines, columns+1] file image ('\\n' in the last column).

    first_line / total_lines: the lines [first_line, first_line + lines) of a file of total_lines lines -- the file is
    generated in chunks of CHUNK lines seeded by their chunk index, so any shard of it is the same bytes no matter
    how the file is cut (bench.py: strong scaling over ranks)."""
    device = torch.device(device)
    if out is None:
        out = torch.empty((lines, columns + 1), dtype=torch.uint8, device=device)
    total = first_line + lines if total_lines is None else total_lines
    tgt = _targets(columns, profile, device)
    w = torch.tensor(CLASS_WEIGHTS, dtype=torch.float32, device=device)
    for ci in range(first_line // CHUNK, (first_line + lines + CHUNK - 1) // CHUNK):
        c0 = ci * CHUNK
        n = min(CHUNK, total - c0)                 # the chunk as the whole file has it (its length decides its random stream)
        lo, hi = max(c0, first_line), min(c0 + n, first_line + lines)
        if lo == c0 and hi == c0 + n:
            _chunk(n, columns, seed, ci, tgt, w, device, out[lo - first_line:hi - first_line])
        else:                                      # a shard boundary inside the chunk: make it whole, keep our part
            tmp = torch.empty((n, columns + 1), dtype=torch.uint8, device=device)
            _chunk(n, columns, seed, ci, tgt, w, device, tmp)
            out[lo - first_line:hi - first_line] = tmp[lo - c0:hi - c0]
    return out

"""Tables for the one-launch weight preparation (tbi_prepare_run / tbi_bn_fold_multi, include/tbi_sm100.h): every per-layer
packing mode of tbi_pack_conv_weights / tbi_pack_convt_weights written as index-map items

    out[out_tap[t] + a*out_a + b*out_b] = S[src_tap_index[t]*src_tap + a*src_a + b] * scale(co)

over one tap's fp32 master matrix S[A][B].  The Engine builds the tables once per set of buffers and the whole model's
compute copies (bf16 K-major forward and data-gradient packs, BatchNorm folded in) are refreshed by two launches per step."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib
from ._lib import BF16, FoldItem, PrepItem


def _item(src, out, gamma, var, ntaps, A, B, src_tap, src_a, out_a, out_b, out_tap, src_tap_index=None, co_is_a=0, co_base=0) -> PrepItem:
    it = PrepItem()
    it.src, it.out, it.gamma, it.var = src, out, gamma, var
    it.ntaps, it.A, it.B, it.co_is_a, it.co_base = ntaps, A, B, co_is_a, co_base
    it.tiles_a, it.tiles_b = (A + 31) // 32, (B + 31) // 32
    it.src_tap, it.src_a, it.out_a, it.out_b = src_tap, src_a, out_a, out_b
    for t in range(ntaps):
        it.src_tap_index[t] = t if src_tap_index is None else src_tap_index[t]
        it.out_tap[t] = out_tap[t]
    return it


def conv_items(L, dt: int, esz: int, mode: int, k: int, groups: int, cin_g: int, cout: int, w_ptr: int, gamma: Optional[int], var: Optional[int],
               out_ptr: int) -> List[PrepItem]:
    """tbi_pack_conv_weights(mode) as items; W is (grouped) HWIO [k*k][cin_g][cout].  Block-diagonal expansion
    (tbi_conv_dense_expand) writes only the diagonal blocks: the destination must have been zero-filled once."""
    nt, cout_g = k * k, cout // groups
    expand = bool(L.tbi_conv_dense_expand(dt, groups, cin_g, cout_g))
    items = []
    if expand:
        cinp = (groups * cin_g + 15) // 16 * 16
        for g in range(groups):
            src = w_ptr + 4 * g * cout_g
            if mode == 0:          # out[co][tap][ci_pad]
                out = out_ptr + esz * ((g * cout_g) * nt * cinp + g * cin_g)
                items.append(_item(src, out, gamma, var, nt, cin_g, cout_g, cin_g * cout, cout, 1, nt * cinp, [t * cinp for t in range(nt)], co_base=g * cout_g))
            else:                  # out[ci_pad][tap'][co]
                out = out_ptr + esz * ((g * cin_g) * nt * cout + g * cout_g)
                items.append(_item(src, out, gamma, var, nt, cin_g, cout_g, cin_g * cout, cout, nt * cout, 1, [(nt - 1 - t) * cout for t in range(nt)], co_base=g * cout_g))
        return items
    if mode == 0:                  # out[co][tap][ci_g]
        return [_item(w_ptr, out_ptr, gamma, var, nt, cin_g, cout, cin_g * cout, cout, 1, nt * cin_g, [t * cin_g for t in range(nt)])]
    for g in range(groups):        # out[g*cin_g+ci][tap'][co_g]
        items.append(_item(w_ptr + 4 * g * cout_g, out_ptr + esz * (g * cin_g * nt * cout_g), gamma, var, nt, cin_g, cout_g, cin_g * cout, cout,
                           nt * cout_g, 1, [(nt - 1 - t) * cout_g for t in range(nt)], co_base=g * cout_g))
    return items


def convt_items(L, esz: int, mode: int, k: int, cin: int, cout: int, cpad: int, w_ptr: int, gamma: Optional[int], var: Optional[int],
                out_ptr: int) -> List[PrepItem]:
    """tbi_pack_convt_weights(mode) as items; W is HWOI [k*k][cout][cin] (A = cout, B = cin).  Padded destinations (cpad > cout,
    head modes 2 and 3) must have been zero-filled once."""
    nt = k * k
    if mode == 0:                  # out[phase][co][t][ci], one item per output-parity phase
        items, off = [], 0
        ky, kx, dy, dx = ((C.c_int * 16)(), (C.c_int * 16)(), (C.c_int * 16)(), (C.c_int * 16)())
        for ph in range(4):
            n = L.tbi_convt_phase_taps(k, ph >> 1, ph & 1, ky, kx, dy, dx)
            base = out_ptr + esz * off * cout * cin
            items.append(_item(w_ptr, base, gamma, var, n, cout, cin, cout * cin, cin, n * cin, 1, [t * cin for t in range(n)],
                               src_tap_index=[ky[t] * k + kx[t] for t in range(n)], co_is_a=1))
            off += n
        return items
    if mode == 1:                  # out[ci][tap][co_pad]
        cp = cpad if cpad > cout else cout
        return [_item(w_ptr, out_ptr, gamma, var, nt, cout, cin, cout * cin, cin, 1, nt * cp, [t * cp for t in range(nt)], co_is_a=1)]
    if mode == 2:                  # out[ci][q], q = tap*cout + co, rows padded to cpad
        return [_item(w_ptr, out_ptr, gamma, var, nt, cout, cin, cout * cin, cin, 1, cpad, [t * cout for t in range(nt)], co_is_a=1)]
    # mode 3: out[q][ci]
    return [_item(w_ptr, out_ptr, gamma, var, nt, cout, cin, cout * cin, cin, cin, 1, [t * cout * cin for t in range(nt)], co_is_a=1)]


class PrepTable:
    """device copies of an item table + its launch parameters"""

    def __init__(self, items: List[PrepItem], device):
        tot = 0
        for it in items:
            it.tile_begin = tot
            tot += it.ntaps * it.tiles_a * it.tiles_b
        self.n, self.total_tiles = len(items), tot
        arr = (PrepItem * len(items))(*items)
        self.dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device) if items else None

    def run(self, L, dt, eps, stream):
        if self.n:
            _lib.check(L.tbi_prepare_run(dt, self.dev.data_ptr(), self.n, self.total_tiles, eps, stream), "prepare_run")


class FoldTable:
    def __init__(self, rows, device):
        """rows: (c, gamma, beta, mean, var, bias, scale, fbias) device pointers (None allowed for the BN four)"""
        items = []
        for (c, g, b, m, v, bias, scale, fbias) in rows:
            it = FoldItem()
            it.c, it.gamma, it.beta, it.mean, it.var, it.bias, it.scale, it.fbias = c, g, b, m, v, bias, scale, fbias
            items.append(it)
        self.n = len(items)
        arr = (FoldItem * len(items))(*items)
        self.dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device) if items else None

    def run(self, L, eps, stream):
        if self.n:
            _lib.check(L.tbi_bn_fold_multi(self.dev.data_ptr(), self.n, eps, stream), "bn_fold_multi")

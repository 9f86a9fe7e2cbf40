"""ctypes binding of libkgwas_b200.so (C ABI declared in include/kgwas_b200.h).

There is no CPU fallback: if the shared library is missing, or a tensor is not a CUDA tensor,
the call raises.  Build the library in-tree with ``python kgwas_b200/csrc/build.py``
(``__graft_entry__.build()`` does it).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkgwas_b200.so")

KGB_NT, KGB_NN, KGB_TN = 0, 1, 2

_lib = None
launches = 0          # number of kernel-launching C-ABI calls made
_prof = None          # bench.py installs a per-launch CUDA-event timer here (None in normal use)


def kernel_launch_count() -> int:
    """Kernels launched by libkgwas_b200 in this process (counted inside the library)."""
    return int(get_lib().kgb_launch_count())


class KgbError(RuntimeError):
    pass


class CsrStruct(C.Structure):
    _fields_ = [("rowptr", C.c_void_p), ("col", C.c_void_p), ("n_rows", C.c_int32), ("seg_len", C.c_int32),
                ("n_hrows", C.c_int32), ("n_hsegs", C.c_int32), ("hrow_id", C.c_void_p),
                ("hrow_segptr", C.c_void_p), ("hseg_hrow", C.c_void_p), ("hseg_order", C.c_void_p),
                ("hrow_grpptr", C.c_void_p), ("n_hgroups", C.c_int32), ("n_edges_hint", C.c_int64),
                ("hitem", C.c_void_p), ("n_hitems", C.c_int32),
                # hub plan (kgb_spmm_hub.cuh; include/kgwas_b200.h)
                ("hub_n", C.c_int32), ("hub_nv", C.c_int32), ("hub_tile_rows", C.c_int32), ("hub_n_tiles", C.c_int32),
                ("hub_n_cta", C.c_int32), ("hub_chunk_cap", C.c_int32), ("hub_n_cols", C.c_int64),
                ("hub_tile_off", C.c_void_p), ("hub_chunks", C.c_void_p), ("hub_row", C.c_void_p),
                ("hub_vptr", C.c_void_p), ("hub_ew", C.c_void_p), ("hitem_tail", C.c_void_p),
                ("n_hitems_tail", C.c_int32), ("n_mid_rows", C.c_int32),
                ("mid_row_id", C.c_void_p)]


_P, _I32, _I64, _F, _SZ = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); must list EVERY symbol include/kgwas_b200.h declares
SIGNATURES = {
    "kgb_version": (C.c_int, []),
    "kgb_sm_arch": (C.c_int, []),
    "kgb_last_error": (C.c_char_p, []),
    "kgb_launch_count": (C.c_longlong, []),
    "kgb_csr_build_workspace_bytes": (_SZ, [_I64, _I64, _I64]),
    "kgb_csr_build": (C.c_int, [_P, _P, _I64, _I64, _I64, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "kgb_csr_heavy_count": (C.c_int, [_P, _I32, _I32, _P, _P, _SZ, _P]),
    "kgb_csr_heavy_workspace_bytes": (_SZ, [_I32]),
    "kgb_csr_heavy_fill": (C.c_int, [_P, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _SZ, _P]),
    "kgb_spmm_scratch_bytes": (_SZ, [_I32, _I32, _I32, _I32]),
    "kgb_spmm_scratch_bytes_csr": (_SZ, [C.POINTER(CsrStruct), _I32]),
    "kgb_spmm": (C.c_int, [C.POINTER(CsrStruct), _P, _P, _P, _P, _I32, _P, _I64, _P, _I64, _I32, _F, _P, _I32, _P, _P,
                           _P, _SZ, _P]),
    "kgb_gat_scratch_bytes": (_SZ, [_I32, _I32]),
    "kgb_gat_alpha": (C.c_int, [C.POINTER(CsrStruct), _P, _P, _I32, _I32, _P, _F, _F, _I32, _P, _SZ, _P]),
    "kgb_sddmm": (C.c_int, [C.POINTER(CsrStruct), _P, _I64, _P, _I64, _I32, _P, _P]),
    "kgb_gat_dsoftmax": (C.c_int, [C.POINTER(CsrStruct), _P, _P, _I32, _I32, _P, _P, _P, _P, _F, _F, _I32, _P, _SZ,
                                   _P]),
    "kgb_gemm_workspace_bytes": (_SZ, [_I32, _I64, _I64, _I64]),
    "kgb_gemm": (C.c_int, [_I32, _P, _I64, _P, _I64, _P, _I64, _I64, _I64, _I64, _F, _F, _P, _I32, _P, _SZ, _P]),
    "kgb_relu_bwd": (C.c_int, [_P, _P, _P, _I64, _P]),
    "kgb_relu_bwd_fused_workspace_bytes": (_SZ, [_I64, _I32]),
    "kgb_relu_bwd_fused": (C.c_int, [_P, _I64, _P, _I64, _P, _P, _F, _P, _I64, _I64, _I32, _P, _P, _SZ, _P]),
    "kgb_wcolsum_workspace_bytes": (_SZ, [_I64, _I32, _I32]),
    "kgb_wcolsum": (C.c_int, [_P, _I64, _P, _I64, _I64, _I32, _I32, _P, _F, _P, _SZ, _P]),
    "kgb_rowdot": (C.c_int, [_P, _I64, _I64, _I32, _I32, _I64, _P, _P, _I64, _P]),
    "kgb_rank_update": (C.c_int, [_P, _I64, _I32, _P, _P, _I64, _I64, _I32, _F, _P]),
    "kgb_permute_f32": (C.c_int, [_P, _P, _P, _I64, _P]),
    "kgb_frontier_workspace_bytes": (_SZ, [_I64]),
    "kgb_frontier_count": (C.c_int, [_P, _P, _I32, _P, C.POINTER(C.c_int32), _P, _SZ, _P]),
    "kgb_frontier_expand": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _P, _P, _P]),
    "kgb_frontier_add": (C.c_int, [_P, _I32, _P, _P, _I32, _P, C.POINTER(C.c_int32), _P, _SZ, _P]),
}


def get_lib():
    """Load the shared library (once).  Raises if it has not been built: no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KgbError(f"{LIB_PATH} not found: the CUDA extension is required (no CPU / PyTorch fallback). "
                           "Build it with `python kgwas_b200/csrc/build.py`.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def _check(rc: int, what: str):
    if rc != 0:
        raise KgbError(f"{what} failed (rc={rc}): {get_lib().kgb_last_error().decode()}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    """Every operand must live on the CURRENT CUDA device: the C ABI launches on the calling thread's device and on
    that device's current stream.  The autograd Functions enter ``torch.cuda.device(x.device)`` (see ``on_device_of``),
    so ``model.to('cuda:1')`` works without ``torch.cuda.set_device``; a direct call with a foreign tensor raises."""
    cur = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise KgbError("kgwas_b200 kernels need CUDA tensors (there is no CPU path); got a CPU tensor")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise KgbError(f"tensor on cuda:{t.device.index} but the current device is cuda:{cur}: wrap the call in "
                           "`with torch.cuda.device(tensor.device)` (the kgwas_b200 modules do this themselves)")


def on_device_of(fn):
    """Decorator for autograd ``forward`` / ``backward`` staticmethods: run with the CUDA device of the first CUDA
    tensor argument current, so that allocations, streams, events and kernel launches all agree with the data."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args):
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args)
                break
        return fn(*args)
    return wrapper


def _f32c(t, name):
    if t.dtype != torch.float32 or t.stride(-1) != 1:
        raise KgbError(f"{name}: expected fp32 with unit inner stride, got {t.dtype} {tuple(t.stride())}")


# ---------------------------------------------------------------------------------------------
# graph bookkeeping
# ---------------------------------------------------------------------------------------------

GAT_LIGHT_MAX = 16      # kLightMax of csrc/kgb_gat.cu: groups up to this size are handled by one thread
SEG_LEN = 128           # edges per heavy-row segment (one warp each); 128 + a 24 MB window measured best on B200
FOLD = 64               # KGB_FOLD: partials folded per level
L2_WINDOW_BYTES = 24 << 20   # gathered-table window the resident warps should share (126 MB L2)


class Csr:
    """CSR of a bipartite edge set + optional transposed CSR + heavy-row segmentation."""

    def __init__(self, rowptr, col, n_rows, n_cols, seg_len=SEG_LEN):
        self.rowptr, self.col, self.n_rows, self.n_cols, self.seg_len = rowptr, col, n_rows, n_cols, seg_len
        self.n_hrows = self.n_hsegs = self.n_hgroups = 0
        self.hrow_id = self.hrow_segptr = self.hseg_hrow = self.hrow_grpptr = None
        self._scratch = {}
        self.hseg_order = None
        self.hitem = None
        self.hub = None
        self.n_mid_rows = -1
        self.mid_row_id = None
        self._build_heavy()
        self._refresh_struct()

    def _refresh_struct(self):
        if self.n_hsegs > 0:
            # one 16-byte record per heavy work item, in scheduling order: (first slot, slot count, segment, heavy row)
            order = self.hseg_order.long() if self.hseg_order is not None else \
                torch.arange(self.n_hsegs, device=self.rowptr.device)
            hr = self.hseg_hrow.long()[order]
            row = self.hrow_id.long()[hr]
            start = self.rowptr.long()[row] + (order - self.hrow_segptr.long()[hr]) * self.seg_len
            length = torch.minimum(self.rowptr.long()[row + 1] - start, torch.full_like(start, self.seg_len))
            self.hitem = torch.stack([start, length, order, hr], dim=1).to(torch.int32).contiguous()
        self.struct = CsrStruct(_ptr(self.rowptr), _ptr(self.col), self.n_rows, self.seg_len, self.n_hrows,
                                self.n_hsegs, _ptr(self.hrow_id), _ptr(self.hrow_segptr), _ptr(self.hseg_hrow),
                                _ptr(self.hseg_order), _ptr(self.hrow_grpptr), self.n_hgroups, int(self.col.numel()),
                                _ptr(self.hitem), self.n_hsegs if self.hitem is not None else 0)
        self.struct.n_mid_rows = self.n_mid_rows
        self.struct.mid_row_id = _ptr(self.mid_row_id)
        hub = self.hub
        if hub is not None:
            st = self.struct
            st.hub_n, st.hub_nv, st.hub_tile_rows, st.hub_n_tiles = hub.n, hub.nv, hub.tile_rows, hub.n_tiles
            st.hub_n_cta, st.hub_chunk_cap, st.hub_n_cols = hub.n_cta, hub.chunk_cap, self.n_cols
            st.hub_tile_off, st.hub_chunks = _ptr(hub.tile_off), _ptr(hub.chunks)
            st.hub_row, st.hub_vptr, st.hub_ew = _ptr(hub.row), _ptr(hub.vptr), _ptr(hub.ew)
            # the pull kernel's work items without the hub rows' segments (same scheduling order)
            keep = ~hub.is_hub_hrow[self.hitem[:, 3].long()]
            hub.hitem_tail = self.hitem[keep].contiguous()
            st.hitem_tail, st.n_hitems_tail = _ptr(hub.hitem_tail), int(hub.hitem_tail.size(0))

    def schedule_for_l2(self, row_bytes: int, window_bytes: int = L2_WINDOW_BYTES):
        """Order the heavy segments by the window of the gathered table they read (rows are column-sorted), so
        that concurrently resident warps gather from the same L2-sized slice instead of sweeping the whole table
        once per hub row.  No-op when the table fits the window."""
        if self.n_hsegs == 0 or self.n_cols * row_bytes <= window_bytes:
            return self
        window_rows = max(1, window_bytes // row_bytes)
        hr = self.hseg_hrow.long()
        seg_in_row = torch.arange(self.n_hsegs, device=hr.device) - self.hrow_segptr.long()[hr]
        first_slot = self.rowptr.long()[self.hrow_id.long()[hr]] + seg_in_row * self.seg_len
        key = torch.div(self.col.long()[first_slot], window_rows, rounding_mode="floor")
        self.hseg_order = torch.argsort(key, stable=True).to(torch.int32)
        self._refresh_struct()
        return self

    @property
    def n_edges(self):
        return self.col.numel()

    def _build_heavy(self):
        if self.n_rows > 0:
            deg = self.rowptr[1:] - self.rowptr[:-1]
            mid = torch.nonzero((deg > GAT_LIGHT_MAX) & (deg <= self.seg_len)).reshape(-1).to(torch.int32)
            self.n_mid_rows = int(mid.numel())
            self.mid_row_id = mid.contiguous() if self.n_mid_rows > 0 else None
        if self.n_rows == 0 or self.col.numel() <= self.seg_len:
            return
        lib = get_lib()
        dev = self.rowptr.device
        ws_bytes = lib.kgb_csr_heavy_workspace_bytes(self.n_rows)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        counts = (C.c_int32 * 2)()
        _check(lib.kgb_csr_heavy_count(_ptr(self.rowptr), self.n_rows, self.seg_len, counts, _ptr(ws), ws_bytes,
                                       _stream()), "kgb_csr_heavy_count")
        self.n_hrows, self.n_hsegs = int(counts[0]), int(counts[1])
        if self.n_hrows == 0:
            return
        self.hrow_id = torch.empty(self.n_hrows, dtype=torch.int32, device=dev)
        self.hrow_segptr = torch.empty(self.n_hrows + 1, dtype=torch.int32, device=dev)
        self.hseg_hrow = torch.empty(self.n_hsegs, dtype=torch.int32, device=dev)
        _check(lib.kgb_csr_heavy_fill(_ptr(self.rowptr), self.n_rows, self.seg_len, self.n_hrows, self.n_hsegs,
                                      _ptr(self.hrow_id), _ptr(self.hrow_segptr), _ptr(self.hseg_hrow), _ptr(ws),
                                      ws_bytes, _stream()), "kgb_csr_heavy_fill")
        nseg = self.hrow_segptr[1:] - self.hrow_segptr[:-1]
        grp = torch.zeros(self.n_hrows + 1, dtype=torch.int32, device=dev)
        grp[1:] = torch.cumsum((nseg + FOLD - 1) // FOLD, 0)
        self.hrow_grpptr = grp
        self.n_hgroups = int(grp[-1].item())

    def scratch(self, h: int):
        """Per-CSR scratch for heavy-row partials and the hub plan's per-CTA partial sets (zeroed once; kernels leave
        the tickets zero)."""
        if self.n_hsegs == 0:
            return None, 0
        # one scratch per launching stream: the ticket counters and partial rows of two gather-reduces over the same CSR
        # that run on different streams (a captured step replaying next to an eager evaluation, model / best_model
        # sharing a plan) must not meet
        key = (h, (self.hub.nv, self.hub.n_cta) if self.hub is not None else None, _stream())
        if key not in self._scratch:
            nbytes = get_lib().kgb_spmm_scratch_bytes_csr(C.byref(self.struct), h)
            self._scratch[key] = torch.zeros(nbytes, dtype=torch.uint8, device=self.rowptr.device)
        s = self._scratch[key]
        return s, s.numel()

    # ------------------------------------------------------------------------------------------------------------
    # hub plan: the heaviest rows reduced from shared-memory tiles (csrc/kgb_spmm_hub.cuh)
    # ------------------------------------------------------------------------------------------------------------
    def build_hub(self, ew: torch.Tensor, h: int, *, min_table_bytes: Optional[int] = None, tile_rows: Optional[int] = None,
                  n_cta: Optional[int] = None, max_slots: Optional[int] = None):
        """Plan the hub-tile path for gather-reduces with the STATIC edge weights ``ew`` (CSR slot order) at feature
        width ``h``.  Idempotent per (ew, h).  Returns True when a plan is attached.

        Chosen when the gathered table is big (>= HUB_MIN_TABLE_BYTES: it cannot live in L2, so every heavy row that
        sweeps it is a DRAM / L2 pass of its own) and the rows are heavy enough to have about one edge per tile."""
        if self.hub is not None and self.hub.ew is ew and self.hub.h == h:
            return True
        self.hub = None
        forced = min_table_bytes is not None                         # explicit request (tests, scratch/bench_hub.py)
        min_table_bytes = HUB_MIN_TABLE_BYTES if min_table_bytes is None else min_table_bytes
        if (h not in (128, 256) or self.n_hsegs == 0 or self.hitem is None or ew is None
                or self.n_cols * h * 4 < min_table_bytes or not (forced or HUB_ENABLED)):
            self._refresh_struct()
            return False
        dev = self.rowptr.device
        T = tile_rows or (128 if h == 128 else 64)
        n_tiles = (self.n_cols + T - 1) // T
        if n_cta is None:
            n_cta = torch.cuda.get_device_properties(dev).multi_processor_count
        n_cta = max(1, min(n_cta, n_tiles))
        W = HUB_WARPS
        rp = self.rowptr.long()
        deg_h = (rp[self.hrow_id.long() + 1] - rp[self.hrow_id.long()])          # degree of every heavy row
        order = torch.argsort(deg_h, descending=True, stable=True)
        deg_sorted = deg_h[order].cpu().numpy()
        # ---- how many hubs: shared memory = slots * row + 2 stages * (tile + chunk); a hub qualifies while it still has
        # about one edge per two tiles; heavy hubs are cut into parts no bigger than half a warp's share
        min_deg = max(self.seg_len + 1, n_tiles // 2)
        budget = max_slots or HUB_MAX_SLOTS[h]
        cand = int((deg_sorted >= min_deg).sum())
        if cand == 0:
            self._refresh_struct()
            return False
        n_try = min(cand, budget)
        while True:
            plan = self._hub_layout(ew, h, T, n_tiles, W, order, deg_sorted, n_try, budget)
            if plan is not None:
                break
            if n_try == 1:
                self._refresh_struct()
                return False
            n_try = max(1, n_try * 3 // 4)
        (n_hub, nv, vptr, hub_hrow, hub_rows, tile_off, chunks, chunk_cap, n_e, slot_sorted, loads) = plan
        is_hub_hrow = torch.zeros(self.n_hrows, dtype=torch.bool, device=dev)
        is_hub_hrow[hub_hrow] = True
        self.hub = _HubPlan(n=n_hub, nv=nv, tile_rows=T, n_tiles=n_tiles, n_cta=n_cta, chunk_cap=chunk_cap, h=h,
                            tile_off=tile_off, chunks=chunks, row=hub_rows.to(torch.int32).contiguous(),
                            vptr=torch.from_numpy(vptr).to(torch.int32).to(dev), ew=ew, is_hub_hrow=is_hub_hrow,
                            n_edges=n_e, slot_csr=slot_sorted, warp_loads=loads)
        self._refresh_struct()
        return True

    def _hub_layout(self, ew, h, T, n_tiles, W, order, deg_sorted, n_try, budget):
        """Slots, warp ownership and per-tile chunks for the ``n_try`` heaviest rows (fewer if cutting the heaviest ones
        into parts needs more than ``budget`` slots); None when the result does not fit one SM's shared memory."""
        import numpy as np
        dev = self.rowptr.device
        rp = self.rowptr.long()
        row_bytes = h * 4
        while True:
            d = deg_sorted[:n_try].astype(np.int64)
            cap = max(1, int(np.ceil(d.sum() / (2.0 * W))))
            parts = np.maximum(1, np.ceil(d / cap)).astype(np.int64)
            if parts.sum() <= budget or n_try == 1:
                break
            n_try -= max(1, int(parts.sum() - budget) // 2)
        if parts.sum() > budget:
            return None
        n_hub, nv = n_try, int(parts.sum())
        hub_hrow = order[:n_hub]                                       # heavy-row slots of the hubs, heaviest first
        vptr = np.zeros(n_hub + 1, dtype=np.int64)
        np.cumsum(parts, out=vptr[1:])
        # ---- slots -> warps: longest-processing-time first
        slot_load = np.repeat(d / parts, parts)
        slot_warp = np.zeros(nv, dtype=np.int64)
        loads = np.zeros(W)
        for sidx in np.argsort(-slot_load, kind="stable"):
            w = int(np.argmin(loads))
            slot_warp[sidx] = w
            loads[w] += slot_load[sidx]
        # ---- every hub edge: (tile, warp, slot, CSR slot) -> sorted records
        hub_rows = self.hrow_id.long()[hub_hrow]                       # row index of every hub
        starts, lens = rp[hub_rows], rp[hub_rows + 1] - rp[hub_rows]
        n_e = int(lens.sum().item())
        hub_of_edge = torch.repeat_interleave(torch.arange(n_hub, device=dev), lens)
        first = torch.cumsum(lens, 0) - lens
        rank = torch.arange(n_e, device=dev) - first[hub_of_edge]      # position inside the hub's row
        slot_csr = starts[hub_of_edge] + rank                          # CSR slot of the edge
        col = self.col.long()[slot_csr]
        tile = torch.div(col, T, rounding_mode="floor")
        parts_t = torch.from_numpy(parts).to(dev)
        vslot = torch.from_numpy(vptr[:-1]).to(dev)[hub_of_edge] + rank % parts_t[hub_of_edge]
        warp = torch.from_numpy(slot_warp).to(dev)[vslot]
        key = (tile * W + warp) * nv + vslot                           # CSR slot order breaks ties (stable sort)
        perm = torch.argsort(key, stable=True)
        key_s, tile_s, vslot_s = key[perm], tile[perm], vslot[perm]
        flush = torch.ones(n_e, dtype=torch.bool, device=dev)
        flush[:-1] = key_s[1:] != key_s[:-1]
        packed = (vslot_s << 8) | (col[perm] - tile_s * T)
        packed = torch.where(flush, packed - (1 << 31), packed).to(torch.int32)
        wbits = ew.contiguous()[slot_csr[perm]].view(torch.int32)
        # ---- chunk layout: per tile [24 x int32 header | records, padded to a multiple of 2]
        cnt_tw = torch.bincount(tile * W + warp, minlength=n_tiles * W).view(n_tiles, W)
        cnt_t = cnt_tw.sum(1)
        rec_pad = (cnt_t + 1) // 2 * 2
        chunk_bytes = HUB_HDR_INTS * 4 + rec_pad * 8
        chunk_cap = (int(chunk_bytes.max().item()) + 32 + 15) // 16 * 16
        if nv * row_bytes + 2 * (T * row_bytes + chunk_cap) + 64 > HUB_SMEM_LIMIT:
            return None
        tile_off = torch.zeros(n_tiles + 1, dtype=torch.int64, device=dev)
        torch.cumsum(chunk_bytes, 0, out=tile_off[1:])
        total_bytes = int(tile_off[-1].item())
        chunks = torch.zeros(total_bytes // 4 + 16, dtype=torch.int32, device=dev)     # + slack: reads past the end
        hdr = torch.zeros((n_tiles, HUB_HDR_INTS), dtype=torch.int32, device=dev)
        hdr[:, 1:W + 1] = torch.cumsum(cnt_tw, 1).to(torch.int32)
        hdr[:, W + 1:] = hdr[:, W:W + 1]
        hdr_pos = (tile_off[:-1] // 4).view(-1, 1) + torch.arange(HUB_HDR_INTS, device=dev).view(1, -1)
        chunks[hdr_pos.reshape(-1)] = hdr.reshape(-1)
        first_rec = torch.cumsum(cnt_t, 0) - cnt_t                     # records before this tile (sorted order)
        rec_in_tile = torch.arange(n_e, device=dev) - first_rec[tile_s]
        pos = tile_off[:-1][tile_s] // 4 + HUB_HDR_INTS + 2 * rec_in_tile
        chunks[pos] = packed
        chunks[pos + 1] = wbits
        return (n_hub, nv, vptr, hub_hrow, hub_rows, tile_off, chunks, chunk_cap, n_e, slot_csr[perm], loads)


class _HubPlan:
    def __init__(self, **kw):
        self.__dict__.update(kw)
        self.hitem_tail = None


HUB_WARPS = 16                      # hub::kWarps
HUB_HDR_INTS = 24                   # hub::kHdrInts
HUB_SMEM_LIMIT = 232448 - 1024      # bytes of dynamic shared memory one CTA may take on sm_100
HUB_MAX_SLOTS = {128: 176, 256: 80}
HUB_MIN_TABLE_BYTES = 64 << 20      # gathered tables smaller than this stay L2-resident: the pull kernel is fine
# Measured on B200 (profiles/r02_hub_tile.md): on kgwas-synth-v1 the tile kernel streams the 401 MB SNP table exactly once
# (434 MB of DRAM reads for 4.02 M hub edges, 137 us) but the hub segments were the CHEAP half of the pull kernel
# (L2-window order keeps them L2-resident): the other 3.98 M edges -- 118 k light and mid rows with uniformly random
# sources -- still cost 331 us on their own, so hub + tail (492 us) loses to the pull kernel alone (444 us).  The path
# is therefore opt-in (KGB_SPMM_HUB=1); it pays on graphs whose tail is small or local.
HUB_ENABLED = os.environ.get("KGB_SPMM_HUB", "0") == "1"


def csr_build(src: torch.Tensor, dst: torch.Tensor, n_src: int, n_dst: int, transposed: bool = True,
              seg_len: int = SEG_LEN, sort_cols: bool = False, presort_key: Optional[torch.Tensor] = None):
    """COO (int64) -> (Csr by dst, eperm, Csr by src | None, t_eperm | None) via kgb_csr_build."""
    _need_cuda(src, dst)
    if src.dtype != torch.int64 or dst.dtype != torch.int64:
        raise KgbError("csr_build: edge indices must be int64")
    src, dst = src.contiguous(), dst.contiguous()
    E, dev = src.numel(), src.device
    lib = get_lib()
    i32 = dict(dtype=torch.int32, device=dev)
    rowptr, col, eperm = torch.empty(n_dst + 1, **i32), torch.empty(E, **i32), torch.empty(E, **i32)
    t_rowptr = t_col = t_eperm = None
    if transposed:
        t_rowptr, t_col, t_eperm = torch.empty(n_src + 1, **i32), torch.empty(E, **i32), torch.empty(E, **i32)
    ws_bytes = lib.kgb_csr_build_workspace_bytes(E, n_src, n_dst)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    global launches
    launches += 1
    _check(lib.kgb_csr_build(_ptr(src), _ptr(dst), E, n_src, n_dst, int(sort_cols),
                             _ptr(presort_key), _ptr(rowptr), _ptr(col), _ptr(eperm),
                             _ptr(t_rowptr), _ptr(t_col), _ptr(t_eperm), _ptr(ws), ws_bytes, _stream()),
           "kgb_csr_build")
    fwd = Csr(rowptr, col, n_dst, n_src, seg_len)
    bwd = Csr(t_rowptr, t_col, n_src, n_dst, seg_len) if transposed else None
    return fwd, eperm, bwd, t_eperm


# ---------------------------------------------------------------------------------------------
# compute
# ---------------------------------------------------------------------------------------------


def spmm(csr: Csr, x: torch.Tensor, y: torch.Tensor, h: int, *, ew=None, wperm=None, ew2=None, rowsum2=None,
         bins: int = 1, beta: float = 0.0, bias=None, relu: bool = False, dot_w=None, dot_out=None):
    """y[i,:h] = act(beta*y[i,:h] + bias + sum_j w_j x[col_j,:h]).  x / y are 2-D views with row strides.
    dot_w [h] / dot_out [n_rows]: also dot_out[i] = <y[i,:h], dot_w> of the finished row (h % 128 == 0, no ew2 /
    wperm: the lean kernel's epilogue)."""
    _need_cuda(x, y, ew, wperm, ew2, rowsum2, bias, dot_w, dot_out)
    _f32c(x, "spmm x"); _f32c(y, "spmm y")
    if csr.n_rows == 0:
        return y
    if csr.n_edges == 0:
        if beta == 0.0:
            y[:, :h].zero_()
        if bias is not None:
            y[:, :h].add_(bias)
        if relu:
            y[:, :h].clamp_(min=0)
        if rowsum2 is not None:
            rowsum2.zero_()
        if dot_w is not None:
            rowdot(y, dot_w, dot_out.view(-1, 1), h, 1, 0)
        return y
    scratch, nbytes = csr.scratch(h)
    global launches
    launches += 1
    if _prof is not None:
        _prof.begin("spmm", csr, h, beta)
    _check(get_lib().kgb_spmm(C.byref(csr.struct), _ptr(ew), _ptr(wperm), _ptr(ew2), _ptr(rowsum2), bins, _ptr(x),
                              x.stride(0), _ptr(y), y.stride(0), h, beta, _ptr(bias), int(relu), _ptr(dot_w),
                              _ptr(dot_out), _ptr(scratch), nbytes, _stream()), "kgb_spmm")
    if _prof is not None:
        _prof.end()
    return y


_gemm_ws = {}


def _workspace(nbytes: int, dev) -> Optional[torch.Tensor]:
    if nbytes == 0:
        return None
    key = (dev.index, torch.cuda.current_stream().cuda_stream)
    ws = _gemm_ws.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 22), dtype=torch.uint8, device=dev)
        _gemm_ws[key] = ws
    return ws


def gemm(layout: int, a: torch.Tensor, b: torch.Tensor, c: torch.Tensor, m: int, n: int, k: int, *,
         alpha: float = 1.0, beta: float = 0.0, bias=None, relu: bool = False):
    """c[m,n] = act(alpha * op(a) op(b) + beta*c + bias); 2-D fp32 views, row strides honoured."""
    _need_cuda(a, b, c, bias)
    _f32c(a, "gemm a"); _f32c(b, "gemm b"); _f32c(c, "gemm c")
    if m == 0 or n == 0:
        return c
    if k == 0:                      # empty contraction (a node type with no rows in this mini-batch): c = act(beta c + b)
        cv = c[:m, :n]
        if beta == 0.0:
            cv.zero_()
        elif beta != 1.0:
            cv.mul_(beta)
        if bias is not None:
            cv.add_(bias)
        if relu:
            cv.clamp_(min=0)
        return c
    lib = get_lib()
    ws_bytes = lib.kgb_gemm_workspace_bytes(layout, m, n, k)
    ws = _workspace(ws_bytes, a.device)
    global launches
    launches += 1
    _check(lib.kgb_gemm(layout, _ptr(a), a.stride(0), _ptr(b), b.stride(0), _ptr(c), c.stride(0), m, n, k,
                        alpha, beta, _ptr(bias), int(relu), _ptr(ws), ws.numel() if ws is not None else 0,
                        _stream()), "kgb_gemm")
    return c


def relu_bwd(dy: torch.Tensor, y: torch.Tensor, out: Optional[torch.Tensor] = None):
    _need_cuda(dy, y)
    dy, y = dy.contiguous(), y.contiguous()
    if out is None:
        out = torch.empty_like(dy)
    global launches
    launches += 1
    _check(get_lib().kgb_relu_bwd(_ptr(dy), _ptr(y), _ptr(out), dy.numel(), _stream()), "kgb_relu_bwd")
    return out


def relu_bwd_fused(g: torch.Tensor, h: int, *, dy=None, y=None, dp=None, wv=None, scale: float = 1.0, sums=None):
    """g[i,:h] = scale * (y[i,:h] > 0) * (dy[i,:h] + dp[i] * wv[:h]);  sums[0] = column sums of g,
    sums[1] = sum_i dp[i] * y[i,:h]  (kgb_relu_bwd_fused)."""
    _need_cuda(g, dy, y, dp, wv, sums)
    for t, name in ((g, "g"), (dy, "dy"), (y, "y")):
        if t is not None:
            _f32c(t, "relu_bwd_fused " + name)
    m = g.size(0)
    if m == 0:                       # a node type without rows in this mini-batch: nothing to mask, zero column sums
        if sums is not None:
            sums.zero_()
        return g
    lib = get_lib()
    ws = None
    if sums is not None:
        if sums.numel() != 2 * h or not sums.is_contiguous():
            raise KgbError("relu_bwd_fused: sums must be a contiguous [2, h] tensor")
        ws = _workspace(lib.kgb_relu_bwd_fused_workspace_bytes(m, h), g.device)
    if dp is not None and (not dp.is_contiguous() or dp.numel() != m):
        raise KgbError("relu_bwd_fused: dp must be a contiguous [m] / [m,1] tensor")
    global launches
    launches += 1
    _check(lib.kgb_relu_bwd_fused(_ptr(dy), dy.stride(0) if dy is not None else 0, _ptr(y),
                                  y.stride(0) if y is not None else 0, _ptr(dp), _ptr(wv), scale, _ptr(g), g.stride(0), m, h,
                                  _ptr(sums), _ptr(ws), ws.numel() if ws is not None else 0, _stream()),
           "kgb_relu_bwd_fused")
    return g


def wcolsum(x: torch.Tensor, h: int, out: torch.Tensor, *, w=None, n_slots: int = 1, beta: float = 0.0):
    """out[s,:h] = beta*out + sum_m w[m,s] x[m,:h]  (w None: plain column sum)."""
    _need_cuda(x, out, w)
    _f32c(x, "wcolsum x")
    m = x.size(0)
    if m == 0:
        if beta == 0.0:
            out.zero_()
        return out
    lib = get_lib()
    nbytes = lib.kgb_wcolsum_workspace_bytes(m, n_slots, h)
    ws = _workspace(nbytes, x.device)
    global launches
    launches += 1
    _check(lib.kgb_wcolsum(_ptr(x), x.stride(0), _ptr(w), w.stride(0) if w is not None else 0, m, n_slots, h,
                           _ptr(out), beta, _ptr(ws), ws.numel(), _stream()), "kgb_wcolsum")
    return out


def rowdot(x: torch.Tensor, v: torch.Tensor, a: torch.Tensor, h: int, n_slots: int, slot_stride: int):
    _need_cuda(x, v, a)
    global launches
    launches += 1
    _check(get_lib().kgb_rowdot(_ptr(x), x.stride(0), x.size(0), n_slots, h, slot_stride, _ptr(v), _ptr(a),
                                a.stride(0), _stream()), "kgb_rowdot")
    return a


def rank_update(a: torch.Tensor, v: torch.Tensor, y: torch.Tensor, h: int, n_slots: int, beta: float):
    _need_cuda(a, v, y)
    global launches
    launches += 1
    _check(get_lib().kgb_rank_update(_ptr(a), a.stride(0), n_slots, _ptr(v), _ptr(y), y.stride(0), y.size(0), h,
                                     beta, _stream()), "kgb_rank_update")
    return y


def permute_f32(w: torch.Tensor, perm: torch.Tensor, out: Optional[torch.Tensor] = None):
    _need_cuda(w, perm)
    if out is None:
        out = torch.empty(perm.numel(), dtype=torch.float32, device=w.device)
    global launches
    launches += 1
    _check(get_lib().kgb_permute_f32(_ptr(w), _ptr(perm), _ptr(out), perm.numel(), _stream()), "kgb_permute_f32")
    return out


# ---------------------------------------------------------------------------------------------
# GAT edge kernels
# ---------------------------------------------------------------------------------------------

ATT_SOFTMAX, ATT_SIGMOID, ATT_RAW = 0, 1, 2


def _gat_scratch(groups: Csr):
    if groups.n_hsegs == 0:
        return None, 0
    key = ("gat", _stream())
    if key not in groups._scratch:
        nbytes = get_lib().kgb_gat_scratch_bytes(groups.n_hrows, groups.n_hsegs)
        groups._scratch[key] = torch.zeros(nbytes, dtype=torch.uint8, device=groups.rowptr.device)
    s = groups._scratch[key]
    return s, s.numel()


def gat_alpha(groups: Csr, a_src, a_dst, n_slots: int, src_is_node: bool, alpha, slope: float, temperature: float,
              mode: int):
    _need_cuda(a_src, a_dst, alpha)
    if groups.n_edges == 0:
        return alpha
    scratch, nbytes = _gat_scratch(groups)
    global launches
    launches += 1
    _check(get_lib().kgb_gat_alpha(C.byref(groups.struct), _ptr(a_src), _ptr(a_dst), n_slots, int(src_is_node),
                                   _ptr(alpha), slope, temperature, mode, _ptr(scratch), nbytes, _stream()),
           "kgb_gat_alpha")
    return alpha


def sddmm(csr: Csr, xrow, x, h: int, out):
    _need_cuda(xrow, x, out)
    _f32c(xrow, "sddmm xrow"); _f32c(x, "sddmm x")
    if csr.n_edges == 0:
        return out
    global launches
    launches += 1
    if _prof is not None:
        _prof.begin("sddmm", csr, h, 0.0)
    _check(get_lib().kgb_sddmm(C.byref(csr.struct), _ptr(xrow), xrow.stride(0), _ptr(x), x.stride(0), h, _ptr(out),
                               _stream()), "kgb_sddmm")
    if _prof is not None:
        _prof.end()
    return out


def gat_dsoftmax(groups: Csr, a_src, a_dst, n_slots: int, src_is_node: bool, alpha, dalpha, du, da_dst, slope: float,
                 temperature: float, mode: int):
    _need_cuda(a_src, a_dst, alpha, dalpha, du, da_dst)
    if groups.n_rows == 0:
        return du, da_dst
    if groups.n_edges == 0:
        da_dst.zero_()
        return du, da_dst
    scratch, nbytes = _gat_scratch(groups)
    global launches
    launches += 1
    _check(get_lib().kgb_gat_dsoftmax(C.byref(groups.struct), _ptr(a_src), _ptr(a_dst), n_slots, int(src_is_node),
                                      _ptr(alpha), _ptr(dalpha), _ptr(du), _ptr(da_dst), slope, temperature, mode,
                                      _ptr(scratch), nbytes, _stream()), "kgb_gat_dsoftmax")
    return du, da_dst


# ---------------------------------------------------------------------------------------------
# full-neighbour frontier expansion (csrc/kgb_sampler.cu)
# ---------------------------------------------------------------------------------------------


def frontier_expand(ptr, col, eperm, frontier):
    """All in-edges of the frontier nodes (frontier order, then edge order): (edge ids int32 [total], sources int32
    [total]).  One host sync (the total has to size the outputs)."""
    _need_cuda(ptr, col, eperm, frontier)
    n_f = int(frontier.numel())
    dev = ptr.device
    i32 = dict(dtype=torch.int32, device=dev)
    if n_f == 0:
        return torch.empty(0, **i32), torch.empty(0, **i32)
    lib = get_lib()
    offsets = torch.empty(n_f + 1, **i32)
    ws = _workspace(lib.kgb_frontier_workspace_bytes(n_f), dev)
    total = C.c_int32(0)
    global launches
    launches += 1
    _check(lib.kgb_frontier_count(_ptr(ptr), _ptr(frontier), n_f, _ptr(offsets), C.byref(total), _ptr(ws), ws.numel(),
                                  _stream()), "kgb_frontier_count")
    t = int(total.value)
    eids, srcs = torch.empty(t, **i32), torch.empty(t, **i32)
    if t:
        launches += 1
        _check(lib.kgb_frontier_expand(_ptr(ptr), _ptr(col), _ptr(eperm), _ptr(frontier), _ptr(offsets), n_f, t,
                                       _ptr(eids), _ptr(srcs), _stream()), "kgb_frontier_expand")
    return eids, srcs


def frontier_add(cand, local, firstpos, count_base: int):
    """Give the not-yet-seen candidates the next local ids in first-occurrence order; returns them (int32 [n_new])."""
    _need_cuda(cand, local, firstpos)
    n = int(cand.numel())
    dev = local.device
    if n == 0:
        return torch.empty(0, dtype=torch.int32, device=dev)
    lib = get_lib()
    new_nodes = torch.empty(min(n, int(local.numel())), dtype=torch.int32, device=dev)
    ws = _workspace(lib.kgb_frontier_workspace_bytes(n), dev)
    n_new = C.c_int32(0)
    global launches
    launches += 1
    _check(lib.kgb_frontier_add(_ptr(cand), n, _ptr(local), _ptr(firstpos), count_base, _ptr(new_nodes), C.byref(n_new),
                                _ptr(ws), ws.numel(), _stream()), "kgb_frontier_add")
    return new_nodes[:int(n_new.value)]

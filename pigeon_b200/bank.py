"""Prototype bank construction for ProtoRefiner: from the reference's on-disk inputs or in-memory objects to the
CSR arrays of `pg_refiner_bank` (include/pigeon_b200.h).

Restates the bank-building cold path of reference models/proto_refiner.py:53-90 (inputs), :257-313 (per-cell
prototype datasets) and :359-378 (prototype = mean of member embeddings, 4-view mean first).
"""
from __future__ import annotations

import json
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch


def _load_indices(index_json) -> list:
    """proto_refiner.py:92-109 — malformed / NaN index lists become []."""
    try:
        return json.loads(index_json)
    except TypeError:
        return []


def bank_from_arrays(proto_cell: np.ndarray, proto_lnglat: np.ndarray, proto_indices: Sequence[Sequence[int]],
                     data_emb: torch.Tensor, data_lnglat: np.ndarray, num_cells: Optional[int] = None,
                     device: str | torch.device = "cpu", proto_count: Optional[Sequence[int]] = None) -> Dict[str, np.ndarray]:
    """Rows (one per prototype, any order) -> CSR bank.  `data_emb` is [N, D] or [N, 4, D] (views averaged first).
    `proto_count` is the CSV `count` column: the reference's single-member shortcut tests it (:243-244, cluster['count'] == 1),
    not the length of the index list, so it is carried as given; omitted, the list lengths are used."""
    proto_cell = np.asarray(proto_cell, np.int64)
    C = int(num_cells if num_cells is not None else proto_cell.max() + 1)
    keep = np.array([len(ix) > 0 for ix in proto_indices], bool)
    # a cell whose FIRST row has no indices is dropped whole by the reference (:304-305); rows inside a kept cell
    # always carry indices in the reference's files, so dropping empty rows is equivalent there.
    order = np.argsort(proto_cell[keep], kind="stable")
    rows = np.nonzero(keep)[0][order]
    cells_sorted = proto_cell[rows]
    counts = np.bincount(cells_sorted, minlength=C)
    cell_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    member_len = np.array([len(proto_indices[r]) for r in rows], np.int64)
    member_off = np.concatenate([[0], np.cumsum(member_len)]).astype(np.int64)
    member_idx = np.fromiter((i for r in rows for i in proto_indices[r]), np.int64, count=int(member_off[-1]))
    emb = torch.as_tensor(data_emb).to(device=device, dtype=torch.float32)
    if emb.is_cuda:
        # GPU builder (pg_bank_build): view mean + one warp per prototype segmented mean
        from ._lib import check, current_stream_ptr, load, ptr
        views = (emb if emb.dim() == 3 else emb.unsqueeze(1)).contiguous()
        n, v, d = views.shape
        mo = torch.as_tensor(member_off, device=emb.device)
        mi = torch.as_tensor(member_idx, device=emb.device)
        if mi.numel() == 0:
            mi = torch.zeros(1, dtype=torch.int64, device=emb.device)
        mean = torch.empty((n, d), dtype=torch.float32, device=emb.device)
        proto_emb = torch.empty((len(rows), d), dtype=torch.float32, device=emb.device)
        check(load().pg_bank_build(ptr(views), n, v, d, ptr(mo), ptr(mi), len(rows), ptr(mean), ptr(proto_emb),
                                   current_stream_ptr()), "pg_bank_build")
        emb = mean
    else:
        # host restatement of the same cold path (CPU-only boxes: fixtures and tests)
        if emb.dim() == 3:
            emb = emb.mean(dim=1)                                        # :370-371
        seg = torch.repeat_interleave(torch.arange(len(rows)), torch.as_tensor(member_len))
        sums = torch.zeros((len(rows), emb.shape[1]), dtype=torch.float32)
        sums.index_add_(0, seg, emb[torch.as_tensor(member_idx)])
        proto_emb = sums / torch.as_tensor(member_len, dtype=torch.float32)[:, None]   # mean over members (:373)
    return dict(cell_off=cell_off, proto_emb=proto_emb.cpu().numpy(),
                proto_lnglat=np.asarray(proto_lnglat, np.float64)[rows].astype(np.float32),
                proto_count=(member_len if proto_count is None else np.asarray(proto_count, np.int64)[rows]).astype(np.int32),
                member_off=member_off, member_idx=member_idx,
                data_emb=emb.cpu().numpy(), data_lnglat=np.asarray(data_lnglat, np.float32))


def bank_from_reference_files(proto_path: str, dataset_path, device: str | torch.device = "cpu") -> Dict[str, np.ndarray]:
    """proto CSV (geocell_idx, cluster, lng, lat, count, indices-json) + HF DatasetDict with a 'train' split holding
    'embedding' and 'labels' columns — the two inputs of the reference constructor (:53-76)."""
    import pandas as pd
    from datasets import DatasetDict, concatenate_datasets
    if isinstance(dataset_path, list):
        if len(dataset_path) > 2:
            raise NotImplementedError('Can\'t concatentate more than 2 datasets.')   # :55 (sic)
        d1, d2 = (DatasetDict.load_from_disk(p) for p in dataset_path)
        train = concatenate_datasets([d1['train'].remove_columns(['labels_climate']),
                                      d2['train'].remove_columns(['labels_climate'])])
    else:
        train = DatasetDict.load_from_disk(dataset_path)['train']
    train = train.with_format('numpy')
    data_emb = torch.from_numpy(np.asarray(train['embedding'], dtype=np.float32))
    data_ll = np.asarray(train['labels'], dtype=np.float32)
    df = pd.read_csv(proto_path)
    idx = [_load_indices(s) for s in df['indices']]
    cells = df['geocell_idx'].astype(int).to_numpy()
    return bank_from_arrays(cells, df[['lng', 'lat']].to_numpy(), idx, data_emb, data_ll,
                            num_cells=int(cells.max()) + 1, device=device, proto_count=df['count'].to_numpy())


def bank_from_proto_rows(cells: Sequence[Optional[Sequence[dict]]], data_emb, data_lnglat) -> Dict[str, np.ndarray]:
    """cells[c] is None or a sequence of rows with keys lng, lat, count, indices, embedding (tensors or numbers),
    i.e. what iterating the reference's per-cell prototype datasets yields.  Prototype embeddings are taken as
    given (they were computed by whoever built `cells`)."""
    emb = torch.as_tensor(np.asarray(data_emb), dtype=torch.float32)
    if emb.dim() == 3:
        emb = emb.mean(dim=1)
    D = emb.shape[1]
    cell_off, pe, pll, cnt, moff, midx = [0], [], [], [], [0], []
    for c in cells:
        if c is not None:
            for r in c:
                pe.append(np.asarray(r["embedding"], np.float32).reshape(D))
                pll.append((np.float32(r["lng"]), np.float32(r["lat"])))
                cnt.append(int(r["count"]))
                midx.extend(int(i) for i in r["indices"])
                moff.append(len(midx))
        cell_off.append(len(pe))
    return dict(cell_off=np.asarray(cell_off, np.int64),
                proto_emb=np.stack(pe).astype(np.float32) if pe else np.zeros((0, D), np.float32),
                proto_lnglat=np.asarray(pll, np.float32).reshape(-1, 2), proto_count=np.asarray(cnt, np.int32),
                member_off=np.asarray(moff, np.int64), member_idx=np.asarray(midx, np.int64),
                data_emb=emb.numpy(), data_lnglat=np.asarray(data_lnglat, np.float32))


def shard_bank(arrays: Dict[str, np.ndarray], rank: int, world: int) -> Dict[str, np.ndarray]:
    """Cell-sharded copy of a CSR bank for `rank` of `world`: geocell c stays iff c % world == rank, every
    other cell becomes an empty range (the reference's `protos[cell] is None`), so a scan over the shard answers "empty cell"
    (-100000) for the pairs it does not own.  Prototype rows, member lists and the member embeddings they point at are
    compacted to the owned cells: the HBM bytes per rank are 1/world of the bank."""
    cell_off = np.asarray(arrays["cell_off"], np.int64)
    C = cell_off.shape[0] - 1
    own = (np.arange(C) % world) == rank
    sizes = np.where(own, cell_off[1:] - cell_off[:-1], 0)
    new_off = np.zeros(C + 1, np.int64)
    np.cumsum(sizes, out=new_off[1:])
    keep = np.concatenate([np.arange(cell_off[c], cell_off[c + 1]) for c in range(C) if own[c]] or
                          [np.zeros(0, np.int64)]).astype(np.int64)
    member_off = np.asarray(arrays["member_off"], np.int64)
    member_idx = np.asarray(arrays["member_idx"], np.int64)
    m_sizes = member_off[keep + 1] - member_off[keep] if keep.size else np.zeros(0, np.int64)
    new_moff = np.zeros(keep.size + 1, np.int64)
    np.cumsum(m_sizes, out=new_moff[1:])
    m_keep = np.concatenate([np.arange(member_off[p], member_off[p + 1]) for p in keep] or
                            [np.zeros(0, np.int64)]).astype(np.int64)
    old_rows = member_idx[m_keep] if m_keep.size else np.zeros(0, np.int64)
    rows, inv = np.unique(old_rows, return_inverse=True) if old_rows.size else (np.zeros(0, np.int64), np.zeros(0, np.int64))
    data_emb = np.asarray(arrays["data_emb"])
    data_ll = np.asarray(arrays["data_lnglat"])
    return dict(cell_off=new_off,
                proto_emb=np.ascontiguousarray(np.asarray(arrays["proto_emb"])[keep]),
                proto_lnglat=np.ascontiguousarray(np.asarray(arrays["proto_lnglat"])[keep]),
                proto_count=np.ascontiguousarray(np.asarray(arrays["proto_count"])[keep]),
                member_off=new_moff, member_idx=inv.astype(np.int64),
                data_emb=np.ascontiguousarray(data_emb[rows]) if rows.size else data_emb[:0],
                data_lnglat=np.ascontiguousarray(data_ll[rows]) if rows.size else data_ll[:0])

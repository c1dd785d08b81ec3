"""Host-side mirror of the reference's operator interface for the retrieval path: same names,
argument meaning, return structure and error behaviour, with the GPU index underneath.

* ``score_results``     <- 05_experiment02.py:200-236
* ``retrieve_colpali``  <- functions.py:884-929
* ``MaxSimClient``      <- the slice of ``qdrant_client.QdrantClient`` those functions and the
  ingestion code use: ``create_collection`` (01_create_context_qdrant.py:208-222), ``upsert``
  (functions.py:865), ``query_points`` (functions.py:894-926, incl. the ``username`` payload filter).

Encoders (``processor`` / ``model``) are the caller's reference PyTorch objects and are used as-is.
"""
from __future__ import annotations

import time
import uuid
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Any, Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

from . import _native as N
from .index import LateInteractionIndex, topk_device
from .scoring import resolve_device

VECTOR_SIZE = N.DIM  # 01_create_context_qdrant.py:70


# ------------------------------------------------------------------------------------------------
# score_results
# ------------------------------------------------------------------------------------------------
_DATASET_INDEX: "OrderedDict[int, tuple]" = OrderedDict()   # id(dataset) -> (fingerprint, index); LRU, bounded
_DATASET_CACHE_MAX = 4


def _embeddings_of(outputs: Any) -> torch.Tensor:
    """colpali_engine models return the embedding tensor; HF ``ColPaliForRetrieval`` returns an
    object with ``.embeddings`` (functions.py:890)."""
    return outputs.embeddings if hasattr(outputs, "embeddings") else outputs


def _dataset_fingerprint(dataset: Sequence[dict]) -> tuple:
    """Cheap content check: length plus storage address, shape and version counter of the embeddings -- replacing
    or editing entries in place (same list object, same length) changes it.  Every entry up to 4096 pages (a few
    hundred microseconds); 256 evenly spaced ones beyond that (call :func:`invalidate_dataset_index` after editing a
    larger dataset in place)."""
    n = len(dataset)
    probe = range(n) if n <= 4096 else sorted({(i * (n - 1)) // 255 for i in range(256)})
    out = [n]
    for i in probe:
        e = dataset[i]["embedding"]
        out.append((e.data_ptr(), tuple(e.shape), str(e.dtype), getattr(e, "_version", 0)))
    return tuple(out)


def invalidate_dataset_index(dataset: Optional[Sequence[dict]] = None) -> None:
    """Drop the cached GPU index of ``dataset`` (or of every dataset) and release its HBM."""
    keys = list(_DATASET_INDEX) if dataset is None else [id(dataset)]
    for k in keys:
        hit = _DATASET_INDEX.pop(k, None)
        if hit is not None:
            hit[1].close()


def index_for_dataset(dataset: Sequence[dict], device=None, dtype: Optional[torch.dtype] = None) -> LateInteractionIndex:
    """Build (once per dataset object) the GPU index over ``entry["embedding"]`` -- this replaces the
    ``torch.stack`` of the whole corpus that the reference repeats on every call (05_experiment02.py:213).
    The cache holds no reference to the dataset, is validated by a content fingerprint, keeps at most
    ``_DATASET_CACHE_MAX`` indexes (least recently used ones are closed) and can be emptied with
    :func:`invalidate_dataset_index`."""
    if len(dataset) == 0:
        raise ValueError("No passages provided")
    key = id(dataset)
    fp = _dataset_fingerprint(dataset)
    hit = _DATASET_INDEX.get(key)
    if hit is not None:
        if hit[0] == fp:
            _DATASET_INDEX.move_to_end(key)
            return hit[1]
        _DATASET_INDEX.pop(key)[1].close()       # same object, different content (or a recycled id)
    embs = [entry["embedding"] for entry in dataset]
    dt = dtype or (embs[0].dtype if embs[0].dtype in (torch.bfloat16, torch.float16, torch.float32) else torch.bfloat16)
    rows = sum(int(e.shape[0]) for e in embs)
    idx = LateInteractionIndex(max(rows, 1), len(embs), dtype=dt, device=device)
    # torch.stack in the reference implies equal lengths; ragged lists get pad_sequence semantics
    idx.add(embs, zero_pad_block=128)
    _DATASET_INDEX[key] = (fp, idx)
    while len(_DATASET_INDEX) > _DATASET_CACHE_MAX:
        _DATASET_INDEX.popitem(last=False)[1][1].close()
    return idx


def load_embedding_cache(path: str) -> List[dict]:
    """Load the reference's page-embedding cache ``data/<retriever>_pdf_emb.pkl`` (05_experiment02.py:391-398):
    a pickled ``list[dict{embedding, doc_id, page_id, file_name}]``.  Feed the result to
    :func:`score_results` / :func:`index_for_dataset` unchanged."""
    import pickle

    with open(path, "rb") as f:
        dataset = pickle.load(f)
    if not isinstance(dataset, list) or (dataset and "embedding" not in dataset[0]):
        raise ValueError(f"{path} is not a page-embedding cache (list of dicts with an 'embedding' entry)")
    return dataset


def convert_embedding_cache(pkl_path: str, out_dir: str, shards: int = 1, device=None,
                            dtype: Optional[torch.dtype] = None) -> dict:
    """Turn the reference's pickle cache (05_experiment02.py:391-398) into the sharded on-disk index format
    (``LateInteractionIndex.save``): page id = position in the list, payload = ``{doc_id, page_id, file_name}``.
    ``shards``: pieces of equal token count, e.g. the number of GPUs that will serve it.  Returns the manifest."""
    dataset = load_embedding_cache(pkl_path)
    if not dataset:
        raise ValueError("No passages provided")
    embs = [e["embedding"] for e in dataset]
    dt = dtype or (embs[0].dtype if embs[0].dtype in (torch.bfloat16, torch.float16, torch.float32) else torch.bfloat16)
    idx = LateInteractionIndex(max(sum(int(e.shape[0]) for e in embs), 1), len(embs), dtype=dt, device=device)
    try:
        idx.add(embs, zero_pad_block=128,
                payloads=[{k: e[k] for k in ("doc_id", "page_id", "file_name") if k in e} for e in dataset])
        return idx.save(out_dir, shards=shards)
    finally:
        idx.close()


def create_document_embeddings(images_per_pdf: dict, model, processor, batch_size: int = 2) -> List[dict]:
    """functions.py:765-809 from already rendered pages (``{file_name: [PIL images]}`` -- what the reference's
    ``convert_pdf_dir_to_images`` returns; PDF rendering itself is out of scope).  Same list-of-dicts result."""
    all_embeddings: List[dict] = []
    for doc_idx, (filename, images) in enumerate(images_per_pdf.items()):
        page_counter = 0
        for i in range(0, len(images), batch_size):
            batch = processor.process_images(images[i:i + batch_size])
            with torch.no_grad():
                inputs = {k: v.to(model.device) for k, v in batch.items()}
                batch_embeddings = _embeddings_of(model(**inputs))
            for embedding in torch.unbind(batch_embeddings.to("cpu")):
                all_embeddings.append({"embedding": embedding, "doc_id": doc_idx, "page_id": page_counter,
                                       "file_name": filename})
                page_counter += 1
    return all_embeddings


def colpali_qdrant(dataset, papers, doi, model, processor, qdrant_client, qdrant_collection, batch_size=4):
    """functions.py:827-873: encode page images in batches and upsert them with the reference's payload.
    With a :class:`MaxSimClient` the embedding tensors go to the index directly (no ``tolist()`` round trip)."""
    direct = isinstance(qdrant_client, MaxSimClient)
    for i in range(0, len(dataset), batch_size):
        batch = dataset[i:i + batch_size]
        images = [item["image"] for item in batch]
        with torch.no_grad():
            batch_images = processor.process_images(images).to(model.device)
            image_embeddings = _embeddings_of(model(**batch_images))
        points = []
        for j, embedding in enumerate(image_embeddings):
            links = [d for paper, d in zip(papers, doi) if paper.split("/")[-1] == batch[j]["filename"]]
            points.append(PointStruct(
                id=str(uuid.uuid4()),
                vector=embedding if direct else embedding.tolist(),
                payload={"document_name": batch[j]["filename"], "document_id": str(uuid.uuid4()),
                         "document_link": links[0] if links else "", "type": "pdf_page", "page_no": batch[j]["page_no"],
                         "ref": "", "caption": "", "img_link": batch[j]["img_link"]}))
        if direct:
            # the in-process index has no transient failures to ride out: an error here (bad vectors, HBM exhausted)
            # would silently lose pages if it were swallowed, so it propagates
            qdrant_client.upsert(collection_name=qdrant_collection, points=points)
            continue
        try:
            qdrant_client.upsert(collection_name=qdrant_collection, points=points)
        except Exception as e:  # a remote Qdrant: the reference skips the batch and carries on (functions.py:866-868)
            print(f"Error during upsert: {e}")
            continue
    print("Indexing complete!")


def score_results(queries: List[str], processor, model, dataset: List[dict], images_per_pdf: dict,
                  top_k: int) -> List[List[dict]]:
    """Retrieve top-k pages per query with late-interaction scoring (05_experiment02.py:200-236)."""
    query_embeddings = processor.process_queries(queries).to(model.device)
    with torch.no_grad():
        query_outputs = _embeddings_of(model(**query_embeddings))
    index = index_for_dataset(dataset, device=query_outputs.device if query_outputs.is_cuda else None)
    scores, ids = index.search(query_outputs, min(top_k, len(dataset)), round_mode="reference")
    retrieved = []
    for q in range(scores.shape[0]):
        results = []
        for s, idx in zip(scores[q].tolist(), ids[q].tolist()):
            if idx < 0:
                continue
            entry = dataset[idx]
            file_name = entry["file_name"]
            page_id = entry["page_id"]
            results.append({
                "doc_id": entry["doc_id"],
                "page_id": page_id,
                "file_name": file_name,
                "image": images_per_pdf[file_name][page_id],
                "score": s,
            })
        retrieved.append(results)
    return retrieved


# ------------------------------------------------------------------------------------------------
# Qdrant-shaped client
# ------------------------------------------------------------------------------------------------
@dataclass
class ScoredPoint:
    id: Any
    score: float
    payload: Optional[dict] = None
    version: int = 0
    vector: Any = None


@dataclass
class QueryResponse:
    points: List[ScoredPoint] = field(default_factory=list)


@dataclass
class PointStruct:
    """Same fields as ``qdrant_client.models.PointStruct`` (functions.py:845-860)."""
    id: Any
    vector: Any
    payload: Optional[dict] = None


@dataclass
class _Collection:
    index: LateInteractionIndex
    point_ids: List[Any]                     # per stored page (position): the caller's point id
    payloads: List[Optional[dict]]
    live: Dict[Any, int] = field(default_factory=dict)   # point id -> position of its current version
    filter_cache: Dict[Any, torch.Tensor] = field(default_factory=dict)


def _filter_conditions(query_filter: Any) -> List[tuple]:
    """Extract ``must=[FieldCondition(key, match=MatchValue(value))]`` pairs from a qdrant
    ``models.Filter`` or an equivalent dict (functions.py:910-918)."""
    if query_filter is None:
        return []
    must = query_filter.get("must") if isinstance(query_filter, dict) else getattr(query_filter, "must", None)
    out = []
    for cond in must or []:
        if isinstance(cond, dict):
            key, match = cond.get("key"), cond.get("match")
            value = match.get("value") if isinstance(match, dict) else getattr(match, "value", None)
        else:
            key = getattr(cond, "key", None)
            value = getattr(getattr(cond, "match", None), "value", None)
        if key is None:
            raise ValueError("only must=[FieldCondition(key=..., match=MatchValue(value=...))] filters are supported")
        out.append((key, value))
    return out


class MaxSimClient:
    """In-process stand-in for the Qdrant server on the multivector MAX_SIM route."""

    def __init__(self, device=None, dtype: torch.dtype = torch.float32, capacity_rows: int = 1_000_000,
                 capacity_pages: int = 2048):
        """``dtype``: storage of the vectors.  float32 (default) is what Qdrant stores for this collection
        (no quantisation configured) and is scored as two bf16 planes with fp32 accumulation;
        ``torch.bfloat16`` halves HBM and doubles throughput at ~1e-2 score error.  The capacities are only the
        initial allocation: the store grows (doubling) when an upsert does not fit."""
        self.device = resolve_device(device)
        self.dtype = dtype
        self.capacity_rows = capacity_rows
        self.capacity_pages = capacity_pages
        self._collections: Dict[str, _Collection] = {}

    # -- schema -----------------------------------------------------------------------------------
    def collection_exists(self, collection_name: str) -> bool:
        return collection_name in self._collections

    def create_collection(self, collection_name: str, vectors_config: Any = None, on_disk_payload: bool = False,
                          capacity_rows: Optional[int] = None, capacity_pages: Optional[int] = None, **_: Any) -> bool:
        """Accepts the reference's ``VectorParams(size=128, distance=COSINE, on_disk=True,
        multivector_config=MultiVectorConfig(comparator=MAX_SIM))`` (01_create_context_qdrant.py:212-221);
        only size 128 / cosine / MAX_SIM is servable."""
        size = getattr(vectors_config, "size", None)
        if isinstance(vectors_config, dict):
            size = vectors_config.get("size")
        if size not in (None, VECTOR_SIZE):
            raise ValueError(f"vector size {size} != {VECTOR_SIZE}")
        if collection_name in self._collections:
            raise ValueError(f"collection {collection_name} already exists")
        idx = LateInteractionIndex(capacity_rows or self.capacity_rows, capacity_pages or self.capacity_pages,
                                   dtype=self.dtype, device=self.device)
        self._collections[collection_name] = _Collection(idx, [], [])
        return True

    def delete_collection(self, collection_name: str) -> bool:
        col = self._collections.pop(collection_name, None)
        if col is not None:
            col.index.close()
        return col is not None

    def count(self, collection_name: str) -> int:
        return len(self._collections[collection_name].live)

    # -- ingestion --------------------------------------------------------------------------------
    def upsert(self, collection_name: str, points: Iterable[Any], **_: Any) -> None:
        """Append points (functions.py:865).  Vectors are cosine-normalised per token like the server
        does for ``Distance.COSINE`` (zero rows stay zero)."""
        col = self._collections[collection_name]
        pages, pids, pays = [], [], []
        for p in points:
            vec = p["vector"] if isinstance(p, dict) else p.vector
            pid = p["id"] if isinstance(p, dict) else p.id
            pay = p.get("payload") if isinstance(p, dict) else p.payload
            v = torch.as_tensor(vec, dtype=torch.float32) if not isinstance(vec, torch.Tensor) else vec.float()
            if v.dim() != 2 or v.shape[1] != VECTOR_SIZE:
                raise ValueError(f"multivector must be [n_tok, {VECTOR_SIZE}], got {tuple(v.shape)}")
            nrm = v.norm(dim=-1, keepdim=True)
            v = torch.where(nrm > 0, v / nrm.clamp_min(1e-30), v)
            pages.append(v.to(self.dtype))
            pids.append(pid if pid is not None else str(uuid.uuid4()))
            pays.append(pay)
        if not pages:
            return
        # a point id seen twice in one call: the last version wins, like consecutive upserts
        last = {pid: j for j, pid in enumerate(pids)}
        keep = sorted(last.values())
        pages, pids, pays = [pages[j] for j in keep], [pids[j] for j in keep], [pays[j] for j in keep]
        idx = col.index
        need_rows = idx.num_rows + sum(int(p.shape[0]) for p in pages)
        need_pages = len(idx) + len(pages)
        cap_rows, cap_pages = idx.capacity
        if need_rows > cap_rows or need_pages > cap_pages:      # grow instead of dropping pages
            idx.reserve(max(need_rows, 2 * cap_rows) if need_rows > cap_rows else cap_rows,
                        max(need_pages, 2 * cap_pages) if need_pages > cap_pages else cap_pages)
        base = len(col.point_ids)
        idx.add(pages, ids=list(range(base, base + len(pages))))
        for j, pid in enumerate(pids):          # upsert semantics: the new version replaces the old one
            old = col.live.get(pid)
            if old is not None:
                idx.tombstone(old)
            col.live[pid] = base + j
        col.point_ids.extend(pids)
        col.payloads.extend(pays)
        col.filter_cache.clear()

    # -- search -----------------------------------------------------------------------------------
    def query_points(self, collection_name: str, query: Any, limit: int = 10, query_filter: Any = None,
                     search_params: Any = None, with_payload: bool = True, **_: Any) -> QueryResponse:
        """MaxSim top-``limit`` (functions.py:894-926).  ``search_params`` (quantisation rescoring) is
        accepted and ignored: scoring here is always exact over the full-precision store."""
        col = self._collections[collection_name]
        if len(col.index) == 0:
            return QueryResponse([])
        q = torch.as_tensor(query, dtype=torch.float32) if not isinstance(query, torch.Tensor) else query.float()
        if q.dim() != 2 or q.shape[1] != VECTOR_SIZE:
            raise ValueError(f"query must be [n_tok, {VECTOR_SIZE}], got {tuple(q.shape)}")
        nrm = q.norm(dim=-1, keepdim=True)
        q = torch.where(nrm > 0, q / nrm.clamp_min(1e-30), q).to(self.dtype)
        k = max(1, min(int(limit), len(col.live), N.MAX_K))
        conds = _filter_conditions(query_filter)
        if not conds:
            scores, ids = col.index.search([q], k)
        else:
            key = tuple(conds)
            masked = col.filter_cache.get(key)
            if masked is None:
                current = np.zeros(len(col.payloads), dtype=bool)
                current[list(col.live.values())] = True          # replaced versions never match
                keep = current & np.asarray([
                    all(((pay or {}).get("metadata", pay or {}).get(kk) == vv) or ((pay or {}).get(kk) == vv)
                        for kk, vv in conds)
                    for pay in col.payloads], dtype=bool)
                masked_np = np.where(keep, np.arange(len(keep), dtype=np.int64), -1)
                masked = torch.from_numpy(masked_np).to(self.device)
                col.filter_cache[key] = masked
            full = col.index.scores([q])
            s_dev, i_dev = topk_device(full, k, ids=masked)
            scores, ids = s_dev.cpu(), i_dev.cpu()
        pts = []
        for s, i in zip(scores[0].tolist(), ids[0].tolist()):
            if i < 0:
                continue
            pts.append(ScoredPoint(id=col.point_ids[i], score=s, payload=col.payloads[i] if with_payload else None))
        return QueryResponse(pts)


def ensure_colpali_collection(client: MaxSimClient, collection_name: str) -> None:
    """01_create_context_qdrant.py:208-222."""
    if client.collection_exists(collection_name):
        return
    client.create_collection(collection_name=collection_name, vectors_config={"size": VECTOR_SIZE}, on_disk_payload=True)


def retrieve_colpali(query, processor, model, qdrant_client, username, colection_name, top_k):
    """functions.py:884-929 with ``qdrant_client`` being a :class:`MaxSimClient` (same keyword
    arguments are forwarded, so a real ``QdrantClient`` also still works)."""
    with torch.no_grad():
        text_embedding = processor.process_queries([query]).to(model.device)
        text_embedding = model(**text_embedding)
    emb = _embeddings_of(text_embedding)[0]
    start_time = time.time()
    if isinstance(qdrant_client, MaxSimClient):
        token_query = emb  # stays on the device: no .tolist() round trip
    else:
        token_query = emb.cpu().float().numpy().tolist()
    if username == "":
        query_result = qdrant_client.query_points(collection_name=colection_name, query=token_query, limit=top_k)
    else:
        query_result = qdrant_client.query_points(
            collection_name=colection_name, query=token_query, limit=top_k,
            query_filter={"must": [{"key": "username", "match": {"value": username}}]})
    print(f"Time taken = {(time.time() - start_time):.3f} s")
    return query_result

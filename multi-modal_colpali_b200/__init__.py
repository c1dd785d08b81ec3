"""B200-native late-interaction (MaxSim) scoring engine -- drop-in for the retrieval hot path of
pkocbek/multi-modal_colpali (``score_multi_vector`` + page-level multivector top-k search).

The directory name contains a hyphen, so import it with
``importlib.import_module("multi-modal_colpali_b200")`` (or ``import mmcolpali_b200`` from the repo root).
"""
from .scoring import (score_multi_vector, plan_queries, clamp_flags, maxsim_scores_device, pack_queries, build_page_store,
                      stream_scores_host_corpus, calibrate_pass_costs)
from .index import LateInteractionIndex, topk_device, merge_topk_device
from .head import project_normalize
from .reference_api import (MaxSimClient, PointStruct, QueryResponse, ScoredPoint, ensure_colpali_collection,
                            retrieve_colpali, score_results, index_for_dataset, load_embedding_cache,
                            create_document_embeddings, colpali_qdrant, convert_embedding_cache,
                            invalidate_dataset_index)
from .sharded import ShardedIndex, Communicator, shard_range, balanced_shard_ranges, assign_shards, gather_candidates
from .batching import QueryBatcher

__all__ = [
    "score_multi_vector", "plan_queries", "clamp_flags", "maxsim_scores_device", "pack_queries", "build_page_store",
    "stream_scores_host_corpus", "calibrate_pass_costs", "convert_embedding_cache", "invalidate_dataset_index", "Communicator", "assign_shards",
    "LateInteractionIndex", "topk_device", "merge_topk_device", "project_normalize",
    "MaxSimClient", "PointStruct", "QueryResponse", "ScoredPoint", "ensure_colpali_collection",
    "retrieve_colpali", "score_results", "index_for_dataset", "load_embedding_cache",
    "create_document_embeddings", "colpali_qdrant",
    "ShardedIndex", "shard_range", "balanced_shard_ranges", "gather_candidates", "QueryBatcher",
]

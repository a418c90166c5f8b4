"""B200-native retrieval-scoring engine: CLIP T2I/T2T similarity scan, knowledge-enhanced
weighted score fusion and top-k / Recall@K / MRR behind the reference's Python API.

    from knowledge_enhanced_multimodal_retrieval_b200 import metrics, fusion, retrieval

mirror `src/clip/eval/metrics.py`, `src/clip/eval/fusion.py` and `src/retrieval.py` of the
reference; `engine` / `index` expose the device-level calls.  Everything on the hot path runs
in libkemr.so (hand-written sm_100a CUDA, C ABI in include/kemr.h); there is no CPU path.
"""
__version__ = "0.1.0"

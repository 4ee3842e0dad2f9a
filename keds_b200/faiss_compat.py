"""`import keds_b200.faiss_compat as faiss` -- the names KEDs uses from the Faiss module
(src/main.py:40,72-83; src/eval_retrieval.py:42,289-296; src/trainer.py:35; src/eval_utils.py)."""
from .index import (  # noqa: F401
    METRIC_INNER_PRODUCT,
    METRIC_L2,
    GpuClonerOptions,
    GpuIndexFlat,
    GpuMultipleClonerOptions,
    IndexFlat,
    IndexFlatIP,
    IndexFlatL2,
    IndexReplicas,
    IndexShards,
    StandardGpuResources,
    get_num_gpus,
    index_cpu_to_all_gpus,
    index_cpu_to_gpu,
    index_gpu_to_cpu,
)

"""Well-known KV keys (reference: tf_yarn/constants.py:1-3)."""
KV_CLUSTER_INSTANCES = "cluster_instances"
KV_EXPERIMENT_FN = "experiment_fn"
KV_TF_SESSION_CONFIG = "tf_session_config"
# additions of the local launcher
KV_GPU_PLACEMENT = "gpu_placement"

"""Root-level shim: `import evaluation` in run.py (run.py:40) resolves to the B200 implementation."""
from gru4rec_b200.evaluation import evaluate_gpu  # noqa: F401

"""Site probe kept for API parity (reference: tf_yarn/_criteo.py:4-9)."""
import os


def is_criteo() -> bool:
    return "CRITEO_ENV" in os.environ

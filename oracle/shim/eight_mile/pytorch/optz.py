import logging  # noqa: F401  (audio8/data.py relies on this star-import)
import torch  # noqa: F401

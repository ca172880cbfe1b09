"""Shim for eight_mile.utils (TEST INFRASTRUCTURE ONLY)."""
from eight_mile_compat import Offsets  # noqa: F401


def str2bool(v):
    return str(v).lower() in ("yes", "true", "t", "1", "y")


def revlut(lut):
    return {v: k for k, v in lut.items()}


def get_num_gpus_multiworker():
    import os
    return int(os.environ.get("WORLD_SIZE", 1))


class Average:
    def __init__(self, name, fmt=":f"):
        self.name, self.fmt = name, fmt
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count

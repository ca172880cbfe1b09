from eight_mile_compat import load_tlm_npz  # noqa: F401

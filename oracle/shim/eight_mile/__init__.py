"""TEST INFRASTRUCTURE ONLY — import shim so the unmodified reference imports (see oracle/README.md)."""

"""ORACLE package — test infrastructure only (see oracle/restatement.py header)."""

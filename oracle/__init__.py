"""ORACLE — test infrastructure only (CPU restatement of the reference head). See head_ref.py."""

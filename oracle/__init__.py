"""ORACLE — test infrastructure only (see oracle/traceobjgrad_oracle.c header). Never imported by juqbox_b200."""
from .jq_oracle import oracle_traceobjgrad, oracle_forward_history, oracle_eval_controls, build, max_threads  # noqa: F401

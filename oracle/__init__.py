"""ORACLE — test infrastructure only (see oracle/cv_front_end.py and oracle/spec.c headers).
Nothing under dynamic_vins_b200/ imports this package."""

#!/bin/bash
# quick GPU check: full GPU test suite + smoke
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -6
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3

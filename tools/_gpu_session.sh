timeout 900 python -m pytest tests -x -q -m gpu -k "api or generation or policy" 2>&1 | tail -3
timeout 300 python tools/dev_step_breakdown.py

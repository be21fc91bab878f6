timeout 300 python -m pytest tests/test_guided.py -m gpu -q -k generate_all 2>&1 | grep -E "^E  |passed|failed|FAILED|Error" | head -30

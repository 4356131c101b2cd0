set -x
python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^E    " | tail -12

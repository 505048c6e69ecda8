# scratch probe used during round 2 (see profiles/round2_summary.md "Multi-GPU"): the GPU suite tail
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | grep -E "passed|failed|^FAILED|^E   .*(assert|Error)" | cut -c1-300 | tail -15

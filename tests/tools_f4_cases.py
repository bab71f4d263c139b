"""The FIR parameterisations of tests/golden/f4_op_grads.npz (tools/make_golden_f4.py:FIR_CASES)."""
FIR_CASES = [('up2',), ('down2',), ('same',), ('generic',), ('crop',)]

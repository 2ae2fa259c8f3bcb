"""Index constants of the DMV score tensors (reference: src/model/torch_struct/dmv.py:7-15).

``dec[b, position, direction, valence, decision]`` and ``attach[b, head, child, valence]``.
"""
NOCHILD = 1
HASCHILD = 0
LEFT = 0
RIGHT = 1
GO = 0
STOP = 1
DIR_NUM = 2
VAL_NUM = 2
DEC_NUM = 2

"""
Utilities mirroring src/vegasflow/utils.py of the reference.
"""
import torch

from vegasflow_b200.configflow import DTYPE, DTYPEINT, float_me, int_me


def consume_array_into_indices(input_arr, indices, result_size):
    """Scatter-add `input_arr[n]` into `result_size` bins chosen by the column
    `indices[n,1]` (utils.py:17-44; the reference builds a one-hot mask, here it
    is a single index_add on the device)."""
    input_arr = float_me(input_arr)
    idx = torch.as_tensor(indices, device=input_arr.device).reshape(-1).to(torch.int64)
    out = torch.zeros(int(result_size), dtype=DTYPE, device=input_arr.device)
    return out.index_add_(0, idx, input_arr)


def py_consume_array_into_indices(input_arr, indices, result_size):
    """utils.py:47-52."""
    return consume_array_into_indices(float_me(input_arr), int_me(indices), int(result_size))


def generate_condition_function(n_mask, condition="and"):
    """Combine `n_mask` boolean masks with and/or and return (mask, indices)
    (utils.py:55-141)."""
    allowed = {"and": torch.logical_and, "or": torch.logical_or}
    if n_mask < 2:
        raise ValueError("At least two masks needed to generate a wrapper")
    if isinstance(condition, str):
        if condition not in allowed:
            raise ValueError(f"Wrong condition {condition}, allowed: {list(allowed)}")
        ops = [allowed[condition]] * (n_mask - 1)
    else:
        if len(condition) != n_mask - 1:
            raise ValueError(f"Wrong number of conditions for {n_mask} masks: {len(condition)}")
        for c in condition:
            if c not in allowed:
                raise ValueError(f"Wrong condition {c}, allowed: {list(allowed)}")
        ops = [allowed[c] for c in condition]

    def condition_to_idx(*masks):
        if len(masks) != n_mask:
            raise ValueError(f"Expected {n_mask} masks, got {len(masks)}")
        res = masks[0]
        for op, m in zip(ops, masks[1:]):
            res = op(res, m)
        return res, torch.nonzero(res).to(DTYPEINT)

    return condition_to_idx

"""Row sharding across ranks: contiguous row ranges per rank, witness replicated, one MIN all-reduce
of the first-unsatisfied GLOBAL row.  Backend-agnostic (NCCL on GPUs; gloo in the CPU tests of this logic)."""

from __future__ import annotations

from typing import Optional, Tuple

SATISFIED = 0x7FFFFFFFFFFFFFFF  # what bp_cs_check_async leaves in device memory when every local row holds


def split_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, exhaustive, near-equal split of [0, n) (units = rows or compression blocks)."""
    assert 0 <= rank < world
    return rank * n // world, (rank + 1) * n // world


def split_units_by_rows(n_units: int, rows_per_unit: int, leading_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of `n_units` equal units (e.g. sha256 compression blocks of `rows_per_unit` rows) that balances ROWS
    when `leading_rows` extra rows (e.g. the boolean rows of the input bits) come before unit 0 and belong to its shard."""
    assert 0 <= rank < world
    total = leading_rows + n_units * rows_per_unit

    def cut(k: int) -> int:  # first unit of rank k
        if k <= 0:
            return 0
        if k >= world:
            return n_units
        return min(n_units, max(0, round((k * total / world - leading_rows) / rows_per_unit)))

    return cut(rank), cut(rank + 1)


def to_global(local_row: int, row_base: int) -> int:
    """Local first-unsatisfied row (-1 = satisfied) -> all-reduce operand."""
    return SATISFIED if local_row < 0 else row_base + local_row


def from_reduced(value: int) -> Optional[int]:
    return None if value == SATISFIED else int(value)


def reduce_first_unsatisfied(result_tensor, world: int):
    """In-place MIN all-reduce of the one-element int64 tensor holding this rank's global first-unsatisfied row."""
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(result_tensor, op=dist.ReduceOp.MIN)
    return result_tensor

"""Queue functions -- mirrors tetris_gymnasium/functional/queue.py (create_bag_queue :20-35, bag_queue_get_next_element :38-67,
create_uniform_queue :71-87, uniform_queue_get_next_element :90-119) on torch CUDA tensors.

Two uses, as in the reference:
  * passed as `create_queue_fn` / `queue_fn` to reset / step they SELECT the device routine `tg_fn_step` runs per env
    (`fn_new_bag`, csrc/tg_fn.cuh) -- the CUDA facade cannot call Python per env;
  * called directly they return what the reference returns, computed by the same device routine (a reset of one throw-away
    env with that key), so a queue built here equals the queue `reset(key)` starts with.
Values come from Philox(rng_key), not from jax.random (threefry): sequences are not JAX-bit-compatible.
Keys are [2] tensors of uint32 values; "advancing" a key increments its second word (the facade's bag counter).
"""
import torch


def _fresh(config, key, queue_fn):
    from ..envs import tetris_fn as F
    from .tetrominoes import TETROMINOES

    key = torch.as_tensor(key).reshape(2)
    _, state, _ = F.reset(TETROMINOES, key, config, create_queue_fn=queue_fn)
    return state.queue[0].clone()


def create_bag_queue(config, key):
    """(queue int32[queue_size] = a permutation of arange(queue_size), queue_index = 0)"""
    return _fresh(config, key, None), torch.zeros((), dtype=torch.int32, device="cuda")


def create_uniform_queue(config, key):
    """(queue int32[queue_size] of draws from [0, queue_size - 1) -- maxval exclusive as in the reference, queue_index = 0)"""
    return _fresh(config, key, "uniform"), torch.zeros((), dtype=torch.int32, device="cuda")


def _next(config, queue, queue_index, key, maker):
    i = int(queue_index)
    if i < config.queue_size:
        return queue[i], queue, torch.as_tensor(i + 1, dtype=torch.int32, device=queue.device), key
    key = torch.as_tensor(key).reshape(2).clone()
    key[1] = (key[1] + 1) & 0xFFFFFFFF
    new, _ = maker(config, key)
    return new[0], new, torch.as_tensor(1, dtype=torch.int32, device=new.device), key


def bag_queue_get_next_element(config, queue, queue_index, key):
    """(element, queue, queue_index, key): the next entry, refilling with a fresh bag once the queue is used up."""
    return _next(config, queue, queue_index, key, create_bag_queue)


def uniform_queue_get_next_element(config, queue, queue_index, key):
    return _next(config, queue, queue_index, key, create_uniform_queue)

"""Queue selectors -- mirrors the names of tetris_gymnasium/functional/queue.py.

The reference passes JAX callables (`create_queue_fn`, `queue_fn`) into reset / step; the CUDA facade cannot call Python per
env, so these functions are SELECTORS: pass them (or the strings "bag" / "uniform") as `create_queue_fn` / `queue_fn` and
`tg_fn_step` runs the matching device routine (`fn_new_bag`, csrc/tg_fn.cuh):

  create_bag_queue / bag_queue_get_next_element           functional/queue.py:20-67   permutations of arange(queue_size)
  create_uniform_queue / uniform_queue_get_next_element   functional/queue.py:71-119  queue_size draws from [0, queue_size - 1)

Values come from Philox(rng_key), not from jax.random (threefry): sequences are not JAX-bit-compatible.
"""


def _selector(name):
    def fn(*args, **kwargs):
        raise TypeError(f"{name} is a queue selector of the CUDA facade (pass it as create_queue_fn / queue_fn), not a callable")
    fn.__name__ = name
    return fn


create_bag_queue = _selector("create_bag_queue")
bag_queue_get_next_element = _selector("bag_queue_get_next_element")
create_uniform_queue = _selector("create_uniform_queue")
uniform_queue_get_next_element = _selector("uniform_queue_get_next_element")

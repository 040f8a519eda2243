from time import time


def timer(func):
    """Wall-clock decorator of the reference (meteotools/tools.py:4-11)."""
    def func_wrapper(*args, **kwargs):
        t0 = time()
        result = func(*args, **kwargs)
        print(f'{func.__name__} cost time: {(time() - t0):.5f} (s).')
        return result
    func_wrapper.__wrapped__ = func
    return func_wrapper

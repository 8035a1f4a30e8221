"""Drop-in for /root/reference/utils/wrapper.py (`@calTimes(logger, msg)`, wrapper.py:12-34).
Unlike the reference the device is synchronised before the clock is read, so the logged
wall time is meaningful for asynchronous CUDA work."""
import functools
import time


def _sync():
    try:
        import torch

        if torch.cuda.is_available():
            torch.cuda.synchronize()
    except Exception:
        pass


def calTimes(logger, msg: str):
    def decorate(func):
        @functools.wraps(func)
        def timed(*args, **kwargs):
            _sync()
            t0 = time.time()
            res = func(*args, **kwargs)
            _sync()
            dt = time.time() - t0
            logger.info((msg + "，" if msg else "") + "耗时：%.2f s" % dt)
            return res

        return timed

    return decorate

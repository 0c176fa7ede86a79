"""6-line stand-in for `diffusers.utils` so /root/reference/ip_adapter/attention_processor.py (its only
third-party import, line 6) can be imported unmodified when generating golden vectors.  Test infrastructure."""
import logging as _logging


def deprecate(*args, **kwargs):
    return None


class logging:  # noqa: N801
    @staticmethod
    def get_logger(name):
        return _logging.getLogger(name)

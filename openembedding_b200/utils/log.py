"""glog-style logging with a role/rank prefix (reference: pico-core common/pico_log.h,
``LogReporter::set_id("WORKER"|"SERVER"|"MASTER", rank)``)."""
import logging
import sys

_role, _rank = "WORKER", 0
_logger = logging.getLogger("openembedding_b200")
if not _logger.handlers:
    h = logging.StreamHandler(sys.stderr)
    h.setFormatter(logging.Formatter("%(levelname).1s%(asctime)s %(role)s %(message)s", "%m%d %H:%M:%S"))
    _logger.addHandler(h)
    _logger.setLevel(logging.INFO)
    _logger.propagate = False


def set_id(role, rank):
    global _role, _rank
    _role, _rank = role, int(rank)


def _extra():
    return {"role": "[%s %d]" % (_role, _rank)}


def info(msg, *a):
    _logger.info(msg, *a, extra=_extra())


def warning(msg, *a):
    _logger.warning(msg, *a, extra=_extra())


def error(msg, *a):
    _logger.error(msg, *a, extra=_extra())


def check(cond, msg="check failed"):
    """SCHECK: log fatal + raise"""
    if not cond:
        _logger.error("FATAL " + msg, extra=_extra())
        raise RuntimeError(msg)

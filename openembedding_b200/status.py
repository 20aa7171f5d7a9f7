"""``Status`` -- the error vocabulary of the parameter-server verbs.

Reference: pico-ps ``Status`` (pico-ps/pico-ps/common/Status.h: OK, INVALID_CONFIG, INVALID_ID,
OOM, TIMEOUT, SERVER_TOO_{NEW,OLD}_CTX[_U], NO_REPLICA, ERROR, FATAL, EMPTY) and the handler retry
state machine that reacts to it (pico-ps/pico-ps/handler/Handler.cpp: retry on TIMEOUT / NO_REPLICA
/ context-version mismatch after refreshing the context from the master).

Here the training data plane is a set of kernels: a failed verb leaves a device-resident error
code (``csrc/cuda/exb_common.cuh: ExbStatus``) that ``CudaEngine.check()`` (blocking) and
``CudaEngine.poll()`` (asynchronous read-back, called every step by ``CudaBackend.tick``) turn into
``StatusError``; the serving client maps transport failures to NO_REPLICA / TIMEOUT and retries
(``serving/client.py``). Context versions: every rank announces a version word into its peers' sync blocks whenever
one of its table slabs moves (alloc, rehash); every plan kernel compares the announced versions with the ones accepted
at the last collective connect (``exb_common.cuh: ctx_check``) and raises ``SERVER_TOO_OLD_CTX`` -- retryable after
``CudaEngine.connect()``. In serving the equivalent is a stale placement record, refreshed from the master tree on
every failed pull.
"""
import enum


class Status(enum.IntEnum):
    OK = 0
    INVALID_CONFIG = 1
    INVALID_ID = 2
    OOM = 3
    TIMEOUT = 4
    SERVER_TOO_NEW_CTX = 5
    SERVER_TOO_OLD_CTX = 6
    SERVER_TOO_NEW_CTX_U = 7
    SERVER_TOO_OLD_CTX_U = 8
    NO_REPLICA = 9
    ERROR = 10
    FATAL = 11
    EMPTY = 12

    @property
    def ok(self):
        return self == Status.OK

    @property
    def retryable(self):
        """what the reference's Handler retries after refreshing its context"""
        return self in (Status.TIMEOUT, Status.NO_REPLICA, Status.SERVER_TOO_NEW_CTX, Status.SERVER_TOO_OLD_CTX,
                        Status.SERVER_TOO_NEW_CTX_U, Status.SERVER_TOO_OLD_CTX_U)


# device error code (ExbStatus) -> Status
ENGINE_STATUS = {
    0: Status.OK,
    1: Status.TIMEOUT,        # grid barrier timeout
    2: Status.TIMEOUT,        # a peer rank did not arrive at the cross-GPU barrier
    3: Status.OOM,            # hash table full
    4: Status.OOM,            # inbox overflow
    5: Status.OOM,            # combine map full
    6: Status.SERVER_TOO_OLD_CTX,   # a peer moved a table slab (rehash / alloc): this rank's mappings are stale -> reconnect
}


class StatusError(RuntimeError):
    def __init__(self, status, message=""):
        self.status = Status(status)
        super().__init__("%s: %s" % (self.status.name, message) if message else self.status.name)


def check(status, message=""):
    if Status(status) != Status.OK:
        raise StatusError(status, message)


def retry(fn, attempts=3, refresh=None):
    """Run ``fn`` and retry retryable StatusErrors (after ``refresh()``), like Handler::wait()."""
    last = None
    for _ in range(max(1, attempts)):
        try:
            return fn()
        except StatusError as e:
            if not e.status.retryable:
                raise
            last = e
            if refresh is not None:
                refresh()
    raise last

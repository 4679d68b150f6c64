"""``mpi4py.MPI`` stand-in over ``torch.distributed`` (one process per GPU).

Covers exactly the calls the reference makes: ``COMM_WORLD`` with
``rank/size``, pickle-style ``allreduce / scatter / alltoall / send / recv / bcast / barrier``,
buffer-style ``Alltoall``, and ``Split_type`` (a single box is one shared-memory node).
"""
import numpy as np

from es_pytorch_b200 import dist as _dist

SUM = 'sum'
ANY_SOURCE = -1
COMM_TYPE_SHARED = 0


class _Float:
    @staticmethod
    def Get_size():
        return 4


FLOAT = _Float()


class Op:
    def __init__(self, fn, commute=True):
        self.fn = fn

    @staticmethod
    def Create(fn, commute=True):
        return Op(fn, commute)


class Comm:
    def __init__(self):
        self._w = _dist.init_from_env()

    rank = property(lambda self: self._w.rank)
    size = property(lambda self: self._w.size)

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def Split_type(self, *a, **k):
        return self

    def Barrier(self):
        self._w.barrier()

    barrier = Barrier

    def bcast(self, obj, root=0):
        return self._w.broadcast_object(obj, root)

    def scatter(self, objs, root=0):
        objs = self._w.broadcast_object(objs, root)
        return objs[self.rank]

    def alltoall(self, objs):
        rows = self._w.allgather_object(list(objs))
        return [rows[src][self.rank] for src in range(self.size)]

    def allreduce(self, obj, op=SUM):
        vals = self._w.allgather_object(obj)
        if isinstance(op, Op):
            acc = vals[0]
            for v in vals[1:]:
                acc = op.fn(acc, v, None)
            return acc
        acc = vals[0]
        for v in vals[1:]:
            acc = acc + v
        return acc

    def Alltoall(self, send, recv):
        send = np.ascontiguousarray(send)
        parts = self._w.allgather_object(send.reshape(self.size, -1))
        np.asarray(recv).reshape(self.size, -1)[...] = np.stack([parts[src][self.rank] for src in range(self.size)])


COMM_WORLD = Comm()

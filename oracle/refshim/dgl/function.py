"""dgl.function builtins used by the reference's sampling path (test infrastructure only)."""


class _USubV:
    def __init__(self, lhs, rhs, out):
        self.lhs, self.rhs, self.out = lhs, rhs, out


class _CopyE:
    def __init__(self, e, out):
        self.e, self.out = e, out


class _Reduce:
    def __init__(self, op, msg, out):
        self.op, self.msg, self.out = op, msg, out


def u_sub_v(lhs, rhs, out):
    return _USubV(lhs, rhs, out)


def copy_e(e, out):
    return _CopyE(e, out)


def sum(msg, out):  # noqa: A001
    return _Reduce('sum', msg, out)


def mean(msg, out):
    return _Reduce('mean', msg, out)

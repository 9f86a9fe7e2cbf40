import math


def glorot(value):
    if value is not None:
        stdv = math.sqrt(6.0 / (value.size(-2) + value.size(-1)))
        value.data.uniform_(-stdv, stdv)


def zeros(value):
    if value is not None:
        value.data.fill_(0.0)

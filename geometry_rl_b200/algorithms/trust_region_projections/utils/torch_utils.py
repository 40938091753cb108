"""utils/torch_utils.py:361-370."""


def inverse_softplus(x):
    return (x.exp() - 1.0).log()

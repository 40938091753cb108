def functional_transform(name):
    def wrapper(cls):
        return cls

    return wrapper

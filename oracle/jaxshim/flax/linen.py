class Module:
    pass

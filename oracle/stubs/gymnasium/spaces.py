class Space:
    pass


class Box(Space):
    pass


class Dict(Space, dict):
    pass

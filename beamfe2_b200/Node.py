"""Node record -- same constructor / attributes as the reference's BeamFE2/Node.py:4-38
(1-based `ID`, `coords=(x, y)`, `x`, `y`, `from_dict`, `set_coords`)."""


class Node(object):
    DOF = ('ux', 'uy', 'rotz')

    def __init__(self, ID=None, coords=()):
        self.ID = ID
        self.coords = coords

    @classmethod
    def from_dict(cls, adict):
        missing = [k for k in ('ID', 'coords') if k not in adict]
        if missing:
            raise Exception('Missing key from dict when creating %s: %s' % (cls.__name__, missing[0]))
        unknown = [k for k in adict if k not in ('ID', 'coords')]
        assert not unknown, 'unknown keys %r' % unknown
        return cls(ID=adict['ID'], coords=adict['coords'])

    def set_coords(self, coords):
        self.coords = coords

    @property
    def x(self):
        return self.coords[0]

    @property
    def y(self):
        return self.coords[1]

    def __repr__(self):
        return 'Node(ID=%d, coords=(%.2f, %.2f)' % (self.ID, self.x, self.y)

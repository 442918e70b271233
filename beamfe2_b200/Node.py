"""Node record of the drop-in API.

Interface contract taken from the reference (BeamFE2/Node.py:4-38): keyword constructor
`Node(ID=<1-based int>, coords=(x, y))`, attributes `ID`, `coords`, read-only `x` / `y`,
`Node.from_dict({'ID': .., 'coords': ..})` that rejects missing or unknown keys, `set_coords`.
Node IDs must be 1..N without gaps: the global DOF position of a node is (ID - 1) * 3 + k
(see beamfe2_b200.plan.position_in_matrix)."""

_FIELDS = ('ID', 'coords')


class _Coordinate(object):
    """Read-only view of one entry of `coords`."""

    def __init__(self, axis):
        self._axis = axis

    def __get__(self, node, owner=None):
        return self if node is None else node.coords[self._axis]


class Node(object):
    DOF = ('ux', 'uy', 'rotz')
    x = _Coordinate(0)
    y = _Coordinate(1)

    def __init__(self, ID=None, coords=()):
        self.ID, self.coords = ID, coords

    @classmethod
    def from_dict(cls, adict):
        for key in _FIELDS:
            if key not in adict:
                raise Exception('Missing key from dict when creating %s: %r' % (cls.__name__, key))
        extra = sorted(set(adict) - set(_FIELDS))
        assert not extra, 'unexpected keys for %s: %s' % (cls.__name__, extra)
        return cls(**{key: adict[key] for key in _FIELDS})

    def set_coords(self, coords):
        self.coords = coords

    def __repr__(self):
        return 'Node(ID=%d, coords=(%.2f, %.2f)' % (self.ID, self.x, self.y)

"""KeypointTensor, host mirror of the API type in
regressor/human_shape/models/body_models/utils.py:123-309: a tensor wrapper carrying keypoint metadata
that callers unwrap through `._t` (demo.py:343-344, evaluation.py:375-376)."""
import torch

_META = ('source', 'keypoint_names', 'connections', 'part_indices', 'part_connections')


def find_joint_kin_chain(joint_id, kinematic_tree):
    chain, cur = [], int(joint_id)
    while cur != -1:
        chain.append(cur)
        cur = int(kinematic_tree[cur])
    return chain


def to_tensor(array, dtype=torch.float32):
    return array.to(dtype=dtype) if torch.is_tensor(array) else torch.tensor(array, dtype=dtype)


class KeypointTensor(object):
    def __init__(self, data, source='smplx', keypoint_names=None, connections=None, part_connections=None,
                 part_indices=None, **kwargs):
        if isinstance(data, KeypointTensor):
            data = data._t
        self._t = torch.as_tensor(data, **kwargs)
        self._source, self._keypoint_names = source, keypoint_names
        self._connections, self._part_indices, self._part_connections = connections, part_indices, part_connections

    def _meta(self):
        return dict(source=self._source, keypoint_names=self._keypoint_names, connections=self._connections,
                    part_indices=self._part_indices, part_connections=self._part_connections)

    @staticmethod
    def from_obj(tensor, obj):
        return KeypointTensor(tensor, **obj._meta())

    source = property(lambda self: self._source)
    keypoint_names = property(lambda self: self._keypoint_names)
    connections = property(lambda self: self._connections)
    part_indices = property(lambda self: self._part_indices)
    part_connections = property(lambda self: self._part_connections)

    def __repr__(self):
        return f'KeypointTensor:\n{self._t}'

    def __getitem__(self, key):
        return self._t[key]

    def __getattr__(self, name):
        # only reached when normal lookup fails: forward to the wrapped tensor
        t = object.__getattribute__(self, '_t')
        attr = getattr(t, name)
        if 'numpy' in name:
            return lambda: t.numpy()
        if callable(attr):
            meta = self._meta()

            def call(*args, **kwargs):
                out = attr(*args, **kwargs)
                return KeypointTensor(out, **meta) if torch.is_tensor(out) else out
            return call
        return attr

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        src = next((a for a in args if isinstance(a, KeypointTensor)), None)
        args = [a._t if isinstance(a, KeypointTensor) else a for a in args]
        ret = func(*args, **kwargs)
        if torch.is_tensor(ret) and src is not None:
            return KeypointTensor(ret, **src._meta())
        return ret

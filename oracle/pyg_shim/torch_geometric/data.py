"""`torch_geometric.data.Data` as the reference uses it (src/layers.py:280,288):
an attribute bag built from a dict whose tensor members follow `.to(device)`."""
import torch


class Data(object):
    @classmethod
    def from_dict(cls, dictionary):
        obj = cls()
        for key, value in dictionary.items():
            setattr(obj, key, value)
        return obj

    def to(self, device):
        def move(v):
            if torch.is_tensor(v):
                return v.to(device)
            if isinstance(v, (list, tuple)):
                return type(v)(move(u) for u in v)
            return v
        for key, value in list(self.__dict__.items()):
            setattr(self, key, move(value))
        return self

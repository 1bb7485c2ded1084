"""Placeholder for `torch_geometric.nn.models.InnerProductDecoder` -- imported by
src/layers.py:2 and only instantiated when MyGAE gets no decoder (never on the TIP path)."""
import torch


class InnerProductDecoder(torch.nn.Module):
    def forward(self, z, edge_index, sigmoid=True):
        value = (z[edge_index[0]] * z[edge_index[1]]).sum(dim=1)
        return torch.sigmoid(value) if sigmoid else value

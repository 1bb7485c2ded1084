"""TEST INFRASTRUCTURE ONLY -- minimal stand-in for torch_geometric 2.0.1.

The reference (NYXFLOWER/TIP) pins `pyg 2.0.1` + `pytorch-scatter 2.0.8`
(environment_tip_gpu.yml:69,79) and neither is installable in this container.
This package restates, on plain torch CPU ops, exactly the slice of that
third-party API the reference's `src/layers.py` touches:

  * `MessagePassing(aggr=...)` / `.propagate(...)`      (src/layers.py:42,78,123,159,202,230)
  * `GCNConv(in, out, cached=True)`                      (src/layers.py:386-387)
  * `Data.from_dict(...).to(device)`                     (src/layers.py:280,288)
  * `InnerProductDecoder` (imported, unused on the path) (src/layers.py:2,258)

It exists so that `oracle/make_golden.py` can import the reference's OWN
`src/layers.py` unmodified and dump golden vectors from the reference's own
module code; only the third-party half is a restatement.  Nothing under
`tip_b200/` may import it.
"""
__version__ = "2.0.1-shim"

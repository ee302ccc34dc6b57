"""Stand-ins for ``torch_geometric`` / ``h5py`` so the reference package can be imported
in THIS container (neither dependency is installed; there is no network).

Test infrastructure only: used by ``make_golden.py`` (fixture generation) and by the
optional ``tests/test_oracle_vs_reference.py`` pinning test.  Never imported by the product.

Restated third-party semantics (torch-geometric >= 2.3, unpinned in the reference's
requirements.txt:5), anchored on the reference's call sites:
  * ``MessagePassing(aggr="add", node_dim=-1).propagate(ei, x=, y=)`` with
    ``message(x_j, y_i)`` (grad_june/infection_networks/base.py:11-13,79-87):
        out = zeros(len(y)).index_add_(0, ei[1], message(x[ei[0]], y[ei[1]]))
    (flow source_to_target: ``_j`` = row 0, ``_i`` = row 1, output sized like the
    ``_i`` side, aggregation = sum in edge order).
  * ``ToUndirected()(data)`` adds ``(dst, "rev_"+rel, src)`` with ``edge_index.flip(0)``
    (grad_june/utils.py:132, june_world_loader/graph_loader.py:38).
  * ``HeteroData``: node stores by name, edge stores by (src, rel, dst) and by rel alone,
    a global store (``data["results"]``), attribute and item access, ``.to(device)``.
"""
import sys
import types

import torch


class _Store:
    def __init__(self, key=None):
        object.__setattr__(self, "_mapping", {})
        object.__setattr__(self, "_key", key)

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        try:
            return self.__dict__["_mapping"][name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self._mapping[name] = value

    def __getitem__(self, name):
        return self._mapping[name]

    def __setitem__(self, name, value):
        self._mapping[name] = value

    def __contains__(self, name):
        return name in self._mapping

    def keys(self):
        return self._mapping.keys()

    def __setstate__(self, state):
        object.__setattr__(self, "_mapping", state.get("_mapping", {}))
        object.__setattr__(self, "_key", state.get("_key"))


class BaseStorage(_Store):
    pass


class NodeStorage(_Store):
    pass


class EdgeStorage(_Store):
    pass


class GlobalStorage(_Store):
    pass


def _move(v, device):
    if torch.is_tensor(v):
        return v.to(device)
    if isinstance(v, dict):
        return {k: _move(x, device) for k, x in v.items()}
    return v


class HeteroData:
    def __init__(self):
        self.__dict__["_global_store"] = GlobalStorage()
        self.__dict__["_node_store_dict"] = {}
        self.__dict__["_edge_store_dict"] = {}

    def __setstate__(self, state):
        self.__dict__.update(state)

    # -- lookup -------------------------------------------------------
    def _find_edge(self, rel):
        for k in self._edge_store_dict:
            if k[1] == rel:
                return k
        return None

    def __getitem__(self, key):
        if isinstance(key, tuple):
            if key not in self._edge_store_dict:
                self._edge_store_dict[key] = EdgeStorage(key)
            return self._edge_store_dict[key]
        k = self._find_edge(key)
        if k is not None:
            return self._edge_store_dict[k]
        if key in self._global_store:
            return self._global_store[key]
        if key not in self._node_store_dict:
            self._node_store_dict[key] = NodeStorage(key)
        return self._node_store_dict[key]

    def __setitem__(self, key, value):
        self._global_store[key] = value

    def __delitem__(self, key):
        if isinstance(key, tuple):
            del self._edge_store_dict[key]
            return
        k = self._find_edge(key)
        if k is not None:
            del self._edge_store_dict[k]
        elif key in self._node_store_dict:
            del self._node_store_dict[key]
        else:
            del self._global_store._mapping[key]

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        gs = self.__dict__["_global_store"]
        if name in gs:
            return gs[name]
        raise AttributeError(name)

    @property
    def edge_types(self):
        return list(self._edge_store_dict.keys())

    def to(self, device):
        for store in list(self._node_store_dict.values()) + list(
            self._edge_store_dict.values()
        ) + [self._global_store]:
            for k, v in list(store._mapping.items()):
                store._mapping[k] = _move(v, device)
        return self


class ToUndirected:
    def __call__(self, data):
        for (src, rel, dst) in list(data._edge_store_dict.keys()):
            if rel.startswith("rev_"):
                continue
            ei = data._edge_store_dict[(src, rel, dst)].edge_index
            data[(dst, "rev_" + rel, src)].edge_index = ei.flip(0)
        return data


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", node_dim=-1, **kw):
        super().__init__()
        assert aggr == "add"

    def propagate(self, edge_index, x, y):
        msg = self.message(x.index_select(-1, edge_index[0]), y.index_select(-1, edge_index[1]))
        out = torch.zeros(y.shape[-1], dtype=msg.dtype, device=msg.device)
        return out.index_add(0, edge_index[1], msg)


def install():
    """Register the stand-ins in sys.modules (idempotent); returns True if the shim is in use,
    False if a real torch_geometric is importable."""
    try:
        import torch_geometric  # noqa: F401
        if not getattr(torch_geometric, "_gj_shim", False):
            return False
        return True
    except ImportError:
        pass
    tg = types.ModuleType("torch_geometric")
    tg._gj_shim = True
    data = types.ModuleType("torch_geometric.data")
    hd = types.ModuleType("torch_geometric.data.hetero_data")
    st = types.ModuleType("torch_geometric.data.storage")
    tr = types.ModuleType("torch_geometric.transforms")
    nn = types.ModuleType("torch_geometric.nn")
    conv = types.ModuleType("torch_geometric.nn.conv")
    data.HeteroData = HeteroData
    hd.HeteroData = HeteroData
    for c in (BaseStorage, NodeStorage, EdgeStorage, GlobalStorage):
        setattr(st, c.__name__, c)
    tr.ToUndirected = ToUndirected
    conv.MessagePassing = MessagePassing
    nn.conv = conv
    nn.MessagePassing = MessagePassing
    tg.data, tg.transforms, tg.nn = data, tr, nn
    data.hetero_data, data.storage = hd, st
    sys.modules.update(
        {
            "torch_geometric": tg,
            "torch_geometric.data": data,
            "torch_geometric.data.hetero_data": hd,
            "torch_geometric.data.storage": st,
            "torch_geometric.transforms": tr,
            "torch_geometric.nn": nn,
            "torch_geometric.nn.conv": conv,
        }
    )
    if "h5py" not in sys.modules:
        try:
            import h5py  # noqa: F401
        except ImportError:
            sys.modules["h5py"] = types.ModuleType("h5py")
    return True


# pytest plugin hook: ``pytest -p _pyg_shim`` installs the stand-ins before collection
install()

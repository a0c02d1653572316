"""CPU stand-in for tinycudann==1.7's torch binding (Encoding only), backed by the
oracle.  TEST INFRASTRUCTURE ONLY.  Mirrors bindings/torch/tinycudann/modules.py:
nn.Module with a single flat ``params`` Parameter, ``n_output_dims``, fp32 output."""
import torch

from oracle import hashgrid as _hg
from oracle.frequency import frequency_encode as _freq


class Encoding(torch.nn.Module):
    def __init__(self, n_input_dims, encoding_config, seed=1337, dtype=None):
        super().__init__()
        self.n_input_dims = n_input_dims
        self.encoding_config = dict(encoding_config)
        self.otype = encoding_config["otype"]
        if self.otype in ("HashGrid", "Grid"):
            self.table = _hg.level_table(
                log2_hashmap_size=encoding_config.get("log2_hashmap_size", 19),
                n_levels=encoding_config.get("n_levels", 16),
                base_resolution=encoding_config.get("base_resolution", 16),
                per_level_scale=float(encoding_config.get("per_level_scale", 2.0)),
                n_features=encoding_config.get("n_features_per_level", 2))
            self.n_output_dims = self.table["n_levels"] * self.table["n_features"]
            self.params = torch.nn.Parameter(_hg.init_params(self.table, seed))
        elif self.otype == "Frequency":
            self.n_frequencies = encoding_config.get("n_frequencies", 12)
            self.n_output_dims = n_input_dims * self.n_frequencies * 2
            self.params = torch.nn.Parameter(torch.zeros(0, dtype=torch.float32))
        elif self.otype == "Identity":
            self.n_output_dims = n_input_dims
            self.params = torch.nn.Parameter(torch.zeros(0, dtype=torch.float32))
        else:
            raise RuntimeError(f"shim: unsupported encoding {self.otype}")

    def forward(self, x):
        x = x.to(torch.float32)
        if self.otype in ("HashGrid", "Grid"):
            return _hg.hashgrid_encode(x, self.params, self.table)
        if self.otype == "Frequency":
            return _freq(x, self.n_frequencies)
        return x

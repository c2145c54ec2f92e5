"""Dataset boundary: npz files -> in-memory list of ``OrderedData`` circuits (reference parser.py:22-125).

Same call surface as upstream -- ``NpzParser(data_dir, circuit_path, label_path, circuit_type, random_shuffle=True,
trainval_split=0.9).get_dataset() -> (train, val)`` -- and the same observable behaviour:
  * two on-disk layouts: AIG keeps everything in ``graphs.npz`` (``edge_index`` [2, E], ``tt_pair_index`` [2, P],
    ``tt_sim``, ``prob``, ``gate``); MIG / XMG / XAG keep ``x``, ``edge_index`` [E, 2] in ``graphs.npz`` and ``tt_dis``,
    ``tt_pair_index`` [P, 2], ``prob`` in ``labels.npz``;
  * seven named circuits and circuits without truth-table pairs are dropped;
  * the parsed set is cached under ``<data_dir>/inmemory/data.pt``; shuffled with ``torch.randperm`` and split 90 / 10.
PyTorch-Geometric is not a dependency: the dataset is a plain indexable list of circuits and batching is
``deepgate.data.collate``.
"""
import os

import numpy as np
import torch

from .data import OrderedData

SKIPPED_CIRCUITS = ("D_FF_0", "register_cc", "D_FF_1", "Main_led_brightness_control_PWM", "ProgramCounter", "TenHertz",
                    "dlatch")                                                     # parser.py:90
_TENSOR_KEYS = ("x", "edge_index", "tt_pair_index", "tt_sim", "forward_level", "forward_index", "backward_level",
                "backward_index", "gate", "prob")


def read_npz_file(filepath):
    """utils/data_utils.py:27-29."""
    return np.load(filepath, allow_pickle=True)


class CircuitDataset(torch.utils.data.Dataset):
    """Indexable list of circuits: ``ds[i]`` is an ``OrderedData``; a slice, index list or index tensor gives a sub-dataset
    (what ``dataset[perm]`` / ``dataset[:cut]`` do on PyG's ``InMemoryDataset``)."""

    def __init__(self, graphs):
        self.graphs = list(graphs)

    def __len__(self):
        return len(self.graphs)

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return CircuitDataset(self.graphs[idx])
        if torch.is_tensor(idx):
            if idx.dim() == 0:
                return self.graphs[int(idx)]
            idx = idx.tolist()
        if isinstance(idx, (list, tuple, np.ndarray)):
            return CircuitDataset([self.graphs[int(i)] for i in idx])
        return self.graphs[int(idx)]

    def __repr__(self):
        return "npz_inmm_dataset({})".format(len(self))


class NpzParser(object):
    """Parse the npz files into an in-memory dataset of ``OrderedData`` circuits."""

    def __init__(self, data_dir, circuit_path, label_path, circuit_type, random_shuffle=True, trainval_split=0.9):
        self.data_dir = data_dir
        self.circuit_type = circuit_type
        dataset = self.inmemory_dataset(data_dir, circuit_path, label_path, circuit_type)
        if random_shuffle:
            dataset = dataset[torch.randperm(len(dataset))]
        cut = int(len(dataset) * trainval_split)
        self.train_dataset = dataset[:cut]
        self.val_dataset = dataset[cut:]

    def get_dataset(self):
        return self.train_dataset, self.val_dataset

    class inmemory_dataset(CircuitDataset):
        def __init__(self, root, circuit_path, label_path, circuit_type, transform=None, pre_transform=None, pre_filter=None):
            self.name = "npz_inmm_dataset"
            self.circuit_type = circuit_type
            self.root = root
            self.circuit_path = circuit_path
            self.label_path = label_path
            if not os.path.exists(self.processed_paths[0]):
                os.makedirs(self.processed_dir, exist_ok=True)
                self.process()
            super().__init__(self._load(self.processed_paths[0]))

        @property
        def raw_dir(self):
            return self.root

        @property
        def processed_dir(self):
            return os.path.join(self.root, "inmemory")

        @property
        def raw_file_names(self):
            return [self.circuit_path, self.label_path]

        @property
        def processed_file_names(self):
            return ["data.pt"]

        @property
        def processed_paths(self):
            return [os.path.join(self.processed_dir, f) for f in self.processed_file_names]

        def process(self):
            if self.circuit_type == "aig":
                from .parser_func import parse_pyg_mlpgate
                tt_key, labels = "tt_sim", None
            else:
                from .parser_func_others import parse_pyg_mlpgate
                tt_key = "tt_dis"
                labels = read_npz_file(self.label_path)["labels"].item()
            circuits = read_npz_file(self.circuit_path)["circuits"].item()
            graphs = []
            for cir_idx, cir_name in enumerate(circuits):
                if cir_name in SKIPPED_CIRCUITS:
                    continue
                print("Parse circuit: {}, {:} / {:} = {:.2f}%".format(cir_name, cir_idx, len(circuits), cir_idx / len(circuits) * 100))
                src = circuits[cir_name] if self.circuit_type == "aig" else labels[cir_name]
                tt_dis, tt_pair_index, prob = src[tt_key], src["tt_pair_index"], src["prob"]
                if len(tt_pair_index) == 0:
                    print("No tt or rc pairs: ", cir_name)
                    continue
                graph = parse_pyg_mlpgate(circuits[cir_name]["x"], circuits[cir_name]["edge_index"], prob, tt_dis, tt_pair_index)
                if self.circuit_type == "aig":
                    graph.gate = torch.as_tensor(np.asarray(circuits[cir_name]["gate"]))
                graph.name = cir_name
                graphs.append(graph)
            torch.save([{k: v for k, v in vars(g).items() if not k.startswith("_")} for g in graphs], self.processed_paths[0])
            print("[INFO] Inmemory dataset save: ", self.processed_paths[0])

        @staticmethod
        def _load(path):
            return [OrderedData(**fields) for fields in torch.load(path, weights_only=False)]

        def __repr__(self):
            return "{}({})".format(self.name, len(self))

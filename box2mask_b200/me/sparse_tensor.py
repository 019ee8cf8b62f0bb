"""SparseTensor and the coordinate manager (host side).

Mirrors ME.SparseTensor as the reference uses it: `ME.SparseTensor(features, coordinates, device=...)`
(/root/reference/models/model.py:43, models/detection_net.py:348,499,503), attributes `.F` / `.C`
(`.C` is mutable in place: `out.C[:,0] = pooling_ids`, models/detection_net.py:347), `out += residual`
(models/resnet.py:80). The coordinate manager owns, per tensor stride, the int32 [N,4] coordinates,
their GPU hash table and the cached kernel maps; every tensor derived from one input shares it.
"""
import torch

from .. import ops


class CoordinateManager:
    """Per-input cache of coordinate levels and kernel maps. Keys are tensor strides (ints)."""

    def __init__(self, coords):
        self.levels = {1: coords}       # stride -> int32 [N,4] on device
        self.tables = {}                # stride -> ops.HashTable
        self.sub_maps = {}              # (stride, ksize) -> sorted ops.KernelMap over the level's rows
        self.stride2 = {}               # fine stride -> (KernelMap down [8, N_coarse], KernelMap up [8, N_fine])
        # unsorted tables kept for the hash-free construction of the next finer level's maps (ops.kernel_map_from_coarse)
        self.parents = {}               # fine stride -> parent row of every fine row, int32 [N_fine]
        self.raw_down = {}              # fine stride -> unsorted child table int32 [8, pitch(N_coarse)]
        self.raw_k3 = {}                # stride -> unsorted 3^3 table of that level

    def coords(self, stride):
        return self.levels[stride]

    def validate(self):
        """Insert the input coordinates into the hash (kept for the kernel maps) and read its status words: one host
        sync. Returns (number of duplicate rows, number of rows outside the packable range)."""
        dups, out_of_range = self.table(1).status.tolist()
        return int(dups), int(out_of_range)

    def deduplicate(self):
        """ME's default quantisation (first occurrence wins): -> (unique_index i64[M] ascending,
        inverse i64[N]) and this manager re-keyed on the M unique coordinates. Rare path (the reference's loaders
        deliver unique coordinates, models/dataloader.py:63-68), plain torch indexing."""
        coords = self.levels[1]
        n = coords.shape[0]
        first = ops.hash_query(self.table(1), coords).long()            # row of the first occurrence of each coordinate
        keep = first == torch.arange(n, device=coords.device)
        unique_index = torch.nonzero(keep).flatten()
        inverse = (torch.cumsum(keep, 0) - 1)[first]
        self.levels = {1: coords[unique_index].contiguous()}
        self.tables, self.sub_maps, self.stride2 = {}, {}, {}
        return unique_index, inverse

    def table(self, stride):
        if stride not in self.tables:
            t = ops.hash_build(self.levels[stride])
            self.tables[stride] = t
        return self.tables[stride]

    def submanifold_map(self, stride, ksize):
        key = (stride, ksize)
        if key not in self.sub_maps:
            coords = self.levels[stride]
            n = coords.shape[0]
            coarse = self.raw_k3.get(2 * stride)
            if ksize in (3, 5) and coarse is not None and stride in self.parents and n > 0:
                # two reads of the coarser level's dense tables per entry instead of a hash probe
                n_coarse = self.levels[2 * stride].shape[0]
                if ksize == 5:      # 125 offsets: never re-ordered, the group masks come out of the same kernel
                    nbr, gmask = ops.kernel_map_from_coarse(coords, stride, ksize, self.parents[stride], coarse,
                                                            self.raw_down[stride], n_coarse, want_gmask=True)
                    self.sub_maps[key] = ops.KernelMap(nbr, None, gmask, ksize ** 3, n)
                    return self.sub_maps[key]
                nbr = ops.kernel_map_from_coarse(coords, stride, ksize, self.parents[stride], coarse,
                                                 self.raw_down[stride], n_coarse)
            else:
                nbr = ops.kernel_map_submanifold(coords, stride, ksize, self.table(stride))
            if ksize == 3 and stride > 1 and (stride // 2) in self.parents:
                self.raw_k3[stride] = nbr          # a finer level exists: its maps will be derived from this table
            self.sub_maps[key] = ops.sort_kernel_map(nbr, n)
        return self.sub_maps[key]

    def stride2_maps(self, fine_stride):
        """Creates the coarse level (2*fine_stride) if needed; returns (KernelMap down, KernelMap up)."""
        if fine_stride not in self.stride2:
            fine = self.levels[fine_stride]
            coarse, parent = ops.downsample_coords(fine, 2 * fine_stride)
            self.levels[2 * fine_stride] = coarse
            nbr_down, nbr_up = ops.kernel_map_stride2(fine, parent, coarse.shape[0], fine_stride)
            self.parents[fine_stride], self.raw_down[fine_stride] = parent, nbr_down
            self.stride2[fine_stride] = (ops.sort_kernel_map(nbr_down, coarse.shape[0]),
                                         ops.sort_kernel_map(nbr_up, fine.shape[0]))
        return self.stride2[fine_stride]


    def prepare(self, n_strided, sub_kernels, stream=None):
        """Build every coordinate level and kernel map a network will ask for, up front. Each strided level costs
        one host sync (its row count); doing them back to back keeps those waits short and leaves the forward /
        backward passes free of host syncs, so the host runs ahead of the GPU.
        n_strided: number of stride-2 levels below the input; sub_kernels: iterable of (tensor_stride, kernel_size).

        stream: build on this (side) CUDA stream instead of the current one — used to construct the maps of the NEXT
        batch while the GPU is still busy with the current step. The consumer calls wait_ready() on its own stream."""
        if stream is not None:
            with torch.cuda.stream(stream):
                self.prepare(n_strided, sub_kernels)
                self._ready = torch.cuda.Event()
                self._ready.record(stream)
            self._side_stream = stream
            return
        s = 1
        for _ in range(int(n_strided)):
            self.stride2_maps(s)
            s *= 2
        # coarsest level first: each finer level's table is derived from the coarser one (only the coarsest is hashed);
        # at equal stride the 3^3 map first (the 5^3 map does not feed anything)
        for stride, ksize in sorted(((int(a), int(b)) for a, b in sub_kernels), key=lambda t: (-t[0], t[1])):
            self.submanifold_map(stride, ksize)
        self.raw_k3.clear()        # construction scratch: the consumers only need the sorted maps
        self.raw_down.clear()

    def _tensors(self):
        for t in self.levels.values():
            yield t
        for tb in self.tables.values():
            yield tb.keys
            yield tb.vals
            yield tb.status
        maps = list(self.sub_maps.values()) + [m for pair in self.stride2.values() for m in pair]
        for m in maps:
            for t in (m.nbr, m.order, m.gmask, m.raw):
                if t is not None:
                    yield t
        for t in self.parents.values():
            yield t

    def wait_ready(self):
        """After a side-stream prepare(): make the current stream wait for the maps and tell the caching allocator
        that they are used on it (they were allocated from the side stream's pool). No-op otherwise."""
        ev = getattr(self, "_ready", None)
        if ev is None:
            return
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        for t in self._tensors():
            t.record_stream(cur)
        self._ready = None


class SparseTensor:
    def __init__(self, features, coordinates=None, device=None, coordinate_manager=None, tensor_stride=1,
                 quantization_mode=None, **_unused):
        if coordinate_manager is None:
            if coordinates is None:
                raise ValueError("SparseTensor needs coordinates or a coordinate_manager")
            if device is None:
                device = features.device
            device = torch.device(device)
            if device.type != "cuda":
                raise ops._lib.B2MError("box2mask_b200 SparseTensor lives on a CUDA device (no CPU path)")
            coords = coordinates.to(device=device, dtype=torch.int32).contiguous()
            if coords.dim() != 2 or coords.shape[1] != 4:
                raise ValueError("coordinates must be [N, 4] = (batch, x, y, z)")
            features = features.to(device)
            if features.shape[0] != coords.shape[0]:
                raise ValueError("features and coordinates disagree on the number of rows")
            coordinate_manager = CoordinateManager(coords)
            tensor_stride = 1
            # ME.SparseTensor inserts the coordinates and quantises: rows outside the representable range are an error,
            # duplicates collapse onto their first occurrence with the features re-indexed (unique_index / inverse).
            self.unique_index = self.inverse_mapping = None
            if coords.shape[0] > 0:
                dups, out_of_range = coordinate_manager.validate()
                if out_of_range:
                    raise RuntimeError("SparseTensor: %d coordinate rows are outside the representable range "
                                       "[-32768, 32767] (batch index < 65535)" % out_of_range)
                if dups:
                    self.unique_index, self.inverse_mapping = coordinate_manager.deduplicate()
                    features = features[self.unique_index]
        self._F = features
        self._manager = coordinate_manager
        self._stride = int(tensor_stride)
        self._C = None
        self._colsum = None   # per-column (sum, sumsq) of F produced by the conv epilogue, consumed by BatchNorm

    # -- reference-visible attributes -------------------------------------------------------------
    @property
    def F(self):
        return self._F

    @property
    def C(self):
        if self._C is None:
            self._C = self._manager.coords(self._stride).clone()   # ME hands out a copy; callers mutate it
        return self._C

    @property
    def coordinate_manager(self):
        return self._manager

    @property
    def tensor_stride(self):
        return [self._stride] * 3

    @property
    def device(self):
        return self._F.device

    @property
    def shape(self):
        return self._F.shape

    def __len__(self):
        return self._F.shape[0]

    def _like(self, features, stride=None):
        return SparseTensor(features, coordinate_manager=self._manager,
                            tensor_stride=self._stride if stride is None else stride)

    def __iadd__(self, other):
        self._check_same_key(other)
        self._F = self._F + other._F
        self._colsum = None
        return self

    def __add__(self, other):
        self._check_same_key(other)
        return self._like(self._F + other._F)

    def _check_same_key(self, other):
        if not isinstance(other, SparseTensor) or other._manager is not self._manager or other._stride != self._stride:
            raise ValueError("SparseTensor arithmetic needs operands on the same coordinate map")

    def __repr__(self):
        return "SparseTensor(rows=%d, channels=%d, tensor_stride=%d, dtype=%s)" % (
            self._F.shape[0], self._F.shape[1], self._stride, self._F.dtype)


TensorField = SparseTensor  # only referenced in a type annotation (models/resnet.py:244)


def cat(*tensors):
    """ME.cat: channel concatenation of tensors on the same coordinate map (models/detection_net.py:286-336)."""
    if len(tensors) == 1 and isinstance(tensors[0], (list, tuple)):
        tensors = tuple(tensors[0])
    first = tensors[0]
    for t in tensors[1:]:
        first._check_same_key(t)
    return first._like(torch.cat([t._F for t in tensors], dim=1))

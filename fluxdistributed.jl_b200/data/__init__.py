"""Data pipeline (reference L3: ``src/imagenet.jl``, ``src/preprocess.jl``,
``src/cifar.jl``, ``Data.toml`` and the Flux-fork prefetching DataLoader)."""
from .cifar import assemble, cifar_key, cifar_minibatch
from .imagenet import fproc, labels, makepaths, minibatch, train_solutions
from .loader import DataLoader, StagedLoader
from .preprocess import (MEAN, STD, center_crop, decode_image, forward, getweights, ground_truth, load_img,
                         make_fname, preprocess, preprocess_u8, resize_smallest_dimension, testimgs, valimgs, weights)
from .registry import BlobTree, RemoteTree, SyntheticTree, dataset, load_registry, open_tree, register
from .synthetic import SyntheticKey, synthetic_minibatch

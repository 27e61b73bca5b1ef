"""root_digger_b200 -- B200-native likelihood engine for RootDigger's hot path.

The product is lib/librdk_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/rdk.h) plus lib/librd_host.so (C++ traversal scheduler and model_t
mirror).  The Python modules are plumbing: build recipes, ctypes bindings,
synthetic inputs and the site-shard planner.
"""
__version__ = "0.1.0"

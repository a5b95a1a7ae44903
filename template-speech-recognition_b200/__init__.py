"""template-speech-recognition_b200 -- B200-native (sm_100a) implementation of the
template_speech_rec data-parallel scoring path: edge spreading, Bernoulli part coding,
part-code spreading, sliding Bernoulli-template LLR detection and Bernoulli-mixture EM.

Reference-signature modules (NumPy in / NumPy out, same names as the reference):
    code_parts, spread_waliji_patches, compute_likelihood_linear_filter,
    bernoulli_mixture, bernoulli_em, edge_signal_proc, estimate_template, test_template
Device-resident batched API: ops, pipeline.  C ABI: include/tsr_b200.h / libtsr_b200.so.

The directory name contains a hyphen, so import it through the alias module at the repo
root (``import tsr_b200``) or importlib.import_module("template-speech-recognition_b200").
There is no CPU fallback: importing is cheap, but every compute call requires the built CUDA
library and a compute-capability-10.x device.
"""
__version__ = "0.1.0"

// TEST / BENCH INFRASTRUCTURE, not product code: the host front ends of the synthetic-state
// generator (hycom_synth_sea_mask, hycom_synth_fill_host) built WITHOUT CUDA, so that the
// reference arm of bench.py (`--impl reference`) can generate its inputs without mapping
// libhycom_tsadvc_b200.so.  Same source as the product's host generator
// (hycom-src_b200/csrc/synth_host.inl on top of synth.h): identical bits.
#include "../hycom-src_b200/csrc/synth_host.inl"

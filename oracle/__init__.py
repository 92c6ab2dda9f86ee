"""CPU oracle of the MSMC-VQ-GAN hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this
package; the product path (msmc-tts_b200/) never does and fails loudly without its CUDA library.

  vq.py          numpy + C (vq_oracle.c) restatement of Quantize / MultiHeadQuantize (bit-exact index oracle)
  ref_modules.py torch-CPU fp32 functional restatement of every module on the path, keyed by the reference's
                 state_dict names (the floating-point oracle and the CPU baseline)
  ref_import.py  imports the UNMODIFIED reference from /root/reference (build container only)
  make_golden.py generates tests/golden/*.pt from the reference itself; the oracle is pinned against them
"""

"""CPU oracle for the decoder hot path -- TEST INFRASTRUCTURE, never imported by the product package.

See `oracle/decoder_oracle.py` for the restatement and `oracle/reference_shim.py` for the loader that
imports the unmodified reference (only where /root/reference exists, i.e. the build container).
"""

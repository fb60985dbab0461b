"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatements (plain PyTorch-eager fp32 / numpy) of the reference's algorithms for the hot
path, used solely as the checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under uwr/ (the product) may import this package.

Pinning: oracle/make_golden.py imports the UNMODIFIED reference (/root/reference, with the
stand-ins under oracle/shims/ for third-party packages that are absent from the image) in the
authoring container and writes tests/golden/*.pt; tests/test_oracle_*.py check every oracle
function against those vectors (and against the live reference when /root/reference exists).
"""

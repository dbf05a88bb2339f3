"""CPU oracle for the ebpmf_log VGA Newton hot path -- TEST INFRASTRUCTURE ONLY.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline and
--impl reference).  See vga_numpy.py / vga_oracle.c headers ("parity unpinned").
"""

"""CPU oracle for the ebpmf_log VGA Newton hot path -- TEST INFRASTRUCTURE ONLY.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline and
--impl reference).  vga_numpy.py / vga_oracle.c: two hand transcriptions of the R code.
mini_r.py / minir_reference.py / standin_flash.py: a restricted R evaluator that executes the
reference's own R source (never GNU R -- see DESIGN.md section 5 for what that does and does not pin).
"""

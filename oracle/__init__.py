"""CPU oracle for the GP hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  The product (gvi_gaussian_process_b200) never imports it and has no CPU fallback.
"""

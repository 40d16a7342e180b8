"""CPU oracle for odgi's path-guided SGD — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product package (odgi_b200) never does.
"""

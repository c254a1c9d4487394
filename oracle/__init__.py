"""ORACLE — CPU restatement of the reference's BigVGAN path.  Test
infrastructure: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it."""

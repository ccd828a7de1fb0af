"""CPU oracle for the D3Q19 alpha/beta time step -- TEST INFRASTRUCTURE ONLY.

* ``oracle.port``  : plain-C restatement of the reference kernels (liblbm_oracle.so)
* ``oracle.ref``   : the reference's own kernel sources compiled as C++ (oracle/_ref)
* ``oracle.multi`` : numpy restatement of CManager decomposition + CController sync

Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU-baseline legs import this
package; the product (turbulent_lbm_multigpu_b200) must never do so.
"""

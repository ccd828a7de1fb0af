"""CConfiguration + the conf.xml schema (reference src/CConfiguration.hpp:22-161, conf.xml).

The reference parses conf.xml with tinyxml2 (a missing submodule); the schema is fixed and
tiny, so the standard-library XML reader is used here.  Tag names are the reference's,
including its ``domian-length`` spelling; ``physics/smagorinsky-constant`` is an optional
addition (default 0 = the reference's plain BGK), old files keep working.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET

TAG_NAME_ROOT = "lbm-configuration"


class CConfiguration:
    def __init__(self, file_name=None):
        # defaults of reference src/main.cpp:100-123
        self.domain_size = (32, 32, 32)
        self.subdomain_num = (1, 1, 1)
        self.domain_length = (0.1, 0.1, 0.1)
        self.gravitation = (0.0, -9.81, 0.0)
        self.viscosity = 0.001308
        self.drivenCavityVelocity = (100.0, 0.0, 0.0, 1.0)
        self.computation_kernel_count = 128
        self.device_nr = 0
        self.do_visualization = False
        self.timestep = -1.0
        self.loops = -1
        self.do_validate = False
        self.lbm_opencl_number_of_registers_list = []
        self.lbm_opencl_number_of_threads_list = []
        self.debug_mode = False
        # additions (not in the reference schema)
        self.smagorinsky_constant = 0.0
        self.precision = "float"
        if file_name is not None:
            self.loadFile(file_name)

    def loadFile(self, file_name):
        try:
            root = ET.parse(file_name).getroot()
        except (OSError, ET.ParseError):
            raise RuntimeError("Loading XML file failed")
        if root.tag != TAG_NAME_ROOT:
            raise RuntimeError("Loading XML file failed")
        self._interpret(root)

    @staticmethod
    def _text(node, path):
        e = node.find(path)
        if e is None or e.text is None:
            raise RuntimeError("conf.xml: missing element %s" % path)
        return e.text.strip()

    def _interpret(self, root):
        t = self._text
        dev = root.find("device")
        self.computation_kernel_count = int(t(dev, "kernel-count"))
        self.device_nr = int(t(dev, "device-number"))
        grid = root.find("grid")
        self.domain_size = tuple(int(t(grid, "domain-size/" + a)) for a in "xyz")
        self.subdomain_num = tuple(int(t(grid, "subdomain-num/" + a)) for a in "xyz")
        self.domain_length = tuple(float(t(grid, "domian-length/" + a)) for a in "xyz")
        phys = root.find("physics")
        self.viscosity = float(t(phys, "viscosity"))
        self.gravitation = tuple(float(t(phys, "gravitation/" + a)) for a in "xyz")
        self.drivenCavityVelocity = tuple(float(t(phys, "cavity-velocity/" + a)) for a in "xyzw")
        sc = phys.find("smagorinsky-constant")
        if sc is not None and sc.text:
            self.smagorinsky_constant = float(sc.text)
        sim = root.find("simulation")
        self.loops = int(t(sim, "loops"))
        self.timestep = float(t(sim, "timestep"))
        self.do_visualization = bool(int(t(sim, "visualization/VTK")))
        self.do_validate = bool(int(t(sim, "validate")))
        pr = sim.find("precision")
        if pr is not None and pr.text:
            self.precision = pr.text.strip()

    def printMe(self):
        print("################\n# CONFIGURATION \n################")
        print("PHYSICS: ")
        print("\t    VISCOSITY: %s" % self.viscosity)
        print("\t  GRAVITATION: %s" % (self.gravitation,))
        print("     CAVITY VEL: %s" % (self.drivenCavityVelocity,))
        print("GRID: ")
        print("\t  DOMAIN_SIZE: %s" % (self.domain_size,))
        print("\tSUBDOMIAN_NUM: %s" % (self.subdomain_num,))
        print("SIMULATION: ")
        print("\t        LOOPS: %s" % self.loops)
        print("\t     TIMESTEP: %s" % self.timestep)
        print("\t          VTK: %d" % self.do_visualization)
        print("\t     VALIDATE: %d" % self.do_validate)
        print("DEVICE: ")
        print("  KERNEL_COUNT: %s" % self.computation_kernel_count)
        print("\t    DEVICE_NR: %s" % self.device_nr)


class ConfigSingleton:
    """Singleton<CConfiguration<T>> (reference src/Singleton.hpp, src/common.h:59)."""
    _instance = None

    @classmethod
    def Instance(cls):
        if cls._instance is None:
            cls._instance = CConfiguration()
        return cls._instance

    @classmethod
    def reset(cls):
        cls._instance = None

"""Makes the reference's import names resolve to this engine:

    import desman_b200.dropin; desman_b200.dropin.install()
    import sampletau                      # -> desman_b200.sampletau
    import desman.HaploSNP_Sampler as h   # -> desman_b200.HaploSNP_Sampler
    import desman.Init_NMFT               # -> desman_b200.Init_NMFT

With only `sampletau` installed (install(classes=False)) the UNMODIFIED reference classes
(HaploSNP_Sampler, Eta_Sampler) run their tau updates on the GPU and, because the module reproduces
the GSL MT19937 stream, produce byte-identical output files.
"""
import sys
import types


def install(classes=True):
    from . import sampletau
    sys.modules["sampletau"] = sampletau
    if classes:
        from . import HaploSNP_Sampler, Init_NMFT, Output_Results, Variant_Filter
        pkg = types.ModuleType("desman")
        pkg.__path__ = []
        for name, mod in (("HaploSNP_Sampler", HaploSNP_Sampler), ("Init_NMFT", Init_NMFT),
                          ("Output_Results", Output_Results), ("Variant_Filter", Variant_Filter)):
            setattr(pkg, name, mod)
            sys.modules["desman." + name] = mod
        sys.modules["desman"] = pkg

/* oracle/ref_stubs: empty stand-in (see optix.h) */

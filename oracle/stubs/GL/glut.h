/* Empty stand-in: DEMO/segment_function.cpp:21-24 includes the GL headers but the
 * hot-path translation units use no GL symbol.  Test infrastructure only. */

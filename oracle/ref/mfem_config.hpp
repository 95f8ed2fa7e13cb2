// Hand-written MFEM configuration header for the serial oracle build.
// TEST INFRASTRUCTURE ONLY (see oracle/README.md). It is passed to the
// vendored MFEM fork through -DMFEM_CONFIG_FILE=... so that nothing has to be
// generated inside (or copied out of) /root/reference. Serial, FP64, OpenMP
// (for the reference's `-d omp` forall path), POSIX clocks.
#ifndef MFEM_CONFIG_HEADER
#define MFEM_CONFIG_HEADER
#define MFEM_VERSION 40701
#define MFEM_VERSION_STRING "4.7.1"
#define MFEM_VERSION_TYPE ((MFEM_VERSION)%2)
#define MFEM_VERSION_TYPE_RELEASE 0
#define MFEM_VERSION_TYPE_DEVELOPMENT 1
#define MFEM_VERSION_MAJOR ((MFEM_VERSION)/10000)
#define MFEM_VERSION_MINOR (((MFEM_VERSION)/100)%100)
#define MFEM_VERSION_PATCH ((MFEM_VERSION)%100)
#define MFEM_SOURCE_DIR "/root/reference/external/mfem-geg"
#define MFEM_INSTALL_DIR "/nonexistent"
#define MFEM_GIT_STRING "(oracle build of the vendored fork)"
#define MFEM_USE_DOUBLE
#define MFEM_USE_OPENMP
#define MFEM_TIMER_TYPE 2
#endif

/*
 * mpi.h shim -- TeaLeaf/comms.h:6 includes <mpi.h> and drivers/remote_halo_driver.c:17 declares an
 * array of MPI_Request; nothing else of MPI is used outside comms.c.  With this backend the ranks of
 * one NVSwitch node talk through comms_b200.cpp (shared memory + NVLink peer stores), so the type is
 * all that is needed.  Build WITHOUT -DNO_MPI (that flag compiles the halo exchange out,
 * remote_halo_driver.c:14,128).
 */
#pragma once
typedef int MPI_Request;

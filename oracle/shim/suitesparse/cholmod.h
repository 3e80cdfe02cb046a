/* TEST INFRASTRUCTURE ONLY: opaque SuiteSparse types so the reference's dense path compiles
 * without SuiteSparse (sparse representation is out of scope). */
#pragma once
struct cholmod_common_struct; typedef struct cholmod_common_struct cholmod_common;
struct cholmod_sparse_struct; typedef struct cholmod_sparse_struct cholmod_sparse;

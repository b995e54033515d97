/*
 * utils.h -- ragged 2-D array helpers, as declared by the reference's include/utils.h:21-46.
 *
 * libdct_cuda does NOT define these four functions: they keep coming from the reference's
 * own, untouched src/utils.c (the "host consumer" side of the boundary).  The header exists
 * so that code written against the reference (`#include <utils.h>` with -Iinclude) compiles
 * against this include directory unchanged.
 */
#ifndef UTILS_H
#define UTILS_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

/* rows x cols doubles: one malloc per row plus one for the row-pointer table, zero-filled */
double **alloc_array(int rows, int cols);
void free_array(double **array, int rows);

/* same shape, int elements */
int **alloc_int_array(int rows, int cols);
void free_int_array(int **array, int rows);

#ifdef __cplusplus
}
#endif

#endif /* UTILS_H */

/* placeholder until the restatement lands in this commit series (see vox_oracle.c) */
#ifndef VOX_ORACLE_H
#define VOX_ORACLE_H
#endif

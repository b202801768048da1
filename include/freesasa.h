/* include/freesasa.h — forwarding header: programs written against the reference's <freesasa.h> (for example its own
 * src/example.c) compile unchanged against the B200-backed host layer.  The supported subset is declared in
 * freesasa_b200_host.h, each function citing the reference lines it mirrors. */
#ifndef FREESASA_H
#define FREESASA_H
#include "freesasa_b200_host.h"
#endif

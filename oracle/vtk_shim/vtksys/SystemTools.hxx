/* stand-in for vtksys/SystemTools.hxx: see ../vtk_standin.h (test infrastructure) */
#include "../vtk_standin.h"

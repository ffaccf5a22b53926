/* stand-in for VTK's vtkDoubleArray.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"

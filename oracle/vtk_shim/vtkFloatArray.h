/* stand-in for VTK's vtkFloatArray.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"

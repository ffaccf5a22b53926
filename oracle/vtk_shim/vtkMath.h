/* stand-in for VTK's vtkMath.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"

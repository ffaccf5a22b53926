/* stand-in for VTK's vtkTransform.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"
